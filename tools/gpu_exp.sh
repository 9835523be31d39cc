#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests_exp.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/tests_exp.log
python tools/layer_table.py train > gpurun_out/layers_train.txt 2>&1; echo "layers train rc=$?"; head -1 gpurun_out/layers_train.txt
python bench.py --workload train --steps 10 --warmup 3 --no-cpu-baseline --detail gpurun_out/train_detail.json > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train rc=$?"; cut -c1-400 gpurun_out/bench_train.json; tail -5 gpurun_out/bench_train.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_train.csv python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_train.log 2>&1; echo "ncu train rc=$?"
