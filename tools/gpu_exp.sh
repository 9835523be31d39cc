#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests_exp.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|Error" gpurun_out/tests_exp.log | tail -12
python bench.py --workload train --steps 20 --warmup 3 --no-cpu-baseline | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_train.csv python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_train.log 2>&1; echo "ncu train rc=$?"
