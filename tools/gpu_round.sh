#!/bin/bash
# Full GPU pass: tests, eval + train bench, ncu launch lists, ncu --set full of the top conv kernel.
mkdir -p gpurun_out
bash tools/gpu_probe.sh tests/test_gpu_confmat.py tests/test_gpu_head_loss.py tests/test_gpu_misc.py tests/test_gpu_conv.py tests/test_gpu_network.py tests/test_gpu_train.py
echo "== tests rc=$?"
python bench.py --steps 10 --warmup 3 --detail gpurun_out/roofline_detail.json > gpurun_out/bench_eval.json 2> gpurun_out/bench_eval.err; echo "bench eval rc=$?"; cat gpurun_out/bench_eval.json
python bench.py --workload train --steps 10 --warmup 3 --detail gpurun_out/train_detail.json > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train rc=$?"; cat gpurun_out/bench_train.json
if [ "$1" != "--no-ncu" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu1.log 2>&1; echo "ncu eval rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_train.csv python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_train.log 2>&1; echo "ncu train rc=$?"
fi
