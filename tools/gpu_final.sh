mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 --detail gpurun_out/r1_eval_classes.json > gpurun_out/r1_bench_eval.json 2> gpurun_out/r1_bench_eval.err; echo "bench eval rc=$?"; cut -c1-160 gpurun_out/r1_bench_eval.json
python bench.py --workload train --steps 20 --warmup 3 --detail gpurun_out/r1_train_classes.json > gpurun_out/r1_bench_train.json 2> gpurun_out/r1_bench_train.err; echo "bench train rc=$?"; cut -c1-160 gpurun_out/r1_bench_train.json
python bench.py --dataset vistas --height 1080 --width 1920 --batch 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_vistas_eval.json 2> gpurun_out/r1_bench_vistas_eval.err; echo "bench vistas eval rc=$?"; cut -c1-160 gpurun_out/r1_bench_vistas_eval.json
python tools/step_timeline.py eval > gpurun_out/r1_timeline_eval.txt 2>&1; head -14 gpurun_out/r1_timeline_eval.txt | tail -12
python -c "import __graft_entry__ as g; g.smoke()"
