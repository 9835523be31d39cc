#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_network.py tests/test_gpu_train.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests_exp.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_exp.log
python tools/layer_table.py train > gpurun_out/layers_train.txt 2>&1; echo "layers train rc=$?"; head -1 gpurun_out/layers_train.txt
python bench.py --workload train --steps 10 --warmup 3 --no-cpu-baseline --detail gpurun_out/train_detail.json > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train rc=$?"; cat gpurun_out/bench_train.json; tail -5 gpurun_out/bench_train.err
