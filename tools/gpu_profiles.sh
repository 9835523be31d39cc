#!/bin/bash
# Round profile pass: plain benches first (exit 0 without ncu), then the ncu launch lists of the same
# commands, DRAM traffic of every conv launch of one step, and --set full captures of the top kernels.
R=${1:-r1}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
python bench.py --steps 20 --warmup 3 --detail gpurun_out/${R}_eval_classes.json > gpurun_out/${R}_bench_eval.json 2> gpurun_out/${R}_bench_eval.err; echo "bench eval rc=$?"; cut -c1-300 gpurun_out/${R}_bench_eval.json
python bench.py --workload train --steps 20 --warmup 3 --detail gpurun_out/${R}_train_classes.json > gpurun_out/${R}_bench_train.json 2> gpurun_out/${R}_bench_train.err; echo "bench train rc=$?"; cut -c1-300 gpurun_out/${R}_bench_train.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err; echo "bench reference rc=$?"; cut -c1-300 gpurun_out/${R}_bench_reference.json
python bench.py --workload train --mixed --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_train_mixed.json 2> gpurun_out/${R}_bench_train_mixed.err; echo "bench train mixed rc=$?"; cut -c1-200 gpurun_out/${R}_bench_train_mixed.json
python bench.py --dataset vistas --height 1080 --width 1920 --batch 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_vistas_eval.json 2> gpurun_out/${R}_bench_vistas_eval.err; echo "bench vistas eval rc=$?"; cut -c1-200 gpurun_out/${R}_bench_vistas_eval.json
python bench.py --workload train --dataset vistas --height 1080 --width 1920 --batch 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_vistas_train.json 2> gpurun_out/${R}_bench_vistas_train.err; echo "bench vistas train rc=$?"; cut -c1-200 gpurun_out/${R}_bench_vistas_train.json
python tools/step_timeline.py train > gpurun_out/${R}_timeline_train.txt 2>&1; head -12 gpurun_out/${R}_timeline_train.txt | tail -10
python tools/step_timeline.py eval > gpurun_out/${R}_timeline_eval.txt 2>&1
python tools/layer_table.py eval > gpurun_out/${R}_layers_eval.txt 2>&1
python tools/layer_table.py train > gpurun_out/${R}_layers_train.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${R}_eval_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --sustained-seconds 0 > gpurun_out/ncu_eval.log 2>&1; echo "ncu eval launches rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/${R}_train_launches.csv python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sustained-seconds 0 > gpurun_out/ncu_train.log 2>&1; echo "ncu train launches rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"conv_igemm|conv_wgrad" -c 400 --csv --log-file gpurun_out/${R}_eval_conv_dram.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --sustained-seconds 0 > gpurun_out/ncu_eval_dram.log 2>&1; echo "ncu eval dram rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"conv_igemm|conv_wgrad|bn_|loss_|head_|confmat|sgdm" -c 3000 --csv --log-file gpurun_out/${R}_train_dram.csv python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sustained-seconds 0 > gpurun_out/ncu_train_dram.log 2>&1; echo "ncu train dram rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"conv_igemm_kernel" --launch-skip 95 --launch-count 8 -o gpurun_out/${R}_eval_igemm256_full -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --sustained-seconds 0 > gpurun_out/ncu_eval_full.log 2>&1; echo "ncu eval full rc=$?"
ncu --set full --clock-control none -k regex:"head_eval|loss_strong|bn_bwd_apply|bn_reduce|bn_apply_rows|conv_wgrad_kernel<256|conv_igemm_kernel<256, __nv_bfloat16, true, 8, true, true" --launch-skip 40 --launch-count 12 -o gpurun_out/${R}_train_bw_full -f python bench.py --workload train --steps 1 --warmup 2 --no-cpu-baseline --no-e2e --sustained-seconds 0 > gpurun_out/ncu_train_full.log 2>&1; echo "ncu train full rc=$?"
ls -la gpurun_out/*.ncu-rep
