#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_head_loss.py tests/test_gpu_train.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests_exp.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|Error" gpurun_out/tests_exp.log | tail -12
python tools/layer_table.py eval 2>&1 | head -2
for m in default all none ty16; do
echo "--- WLSEG_LOSS_COLS=$m"
if [ $m = default ]; then unset WLSEG_LOSS_COLS; else export WLSEG_LOSS_COLS=$m; fi
python bench.py --workload train --mixed --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --detail gpurun_out/mixed_detail.json | cut -c1-150; grep -A2 loss_fwd gpurun_out/mixed_detail.json | tail -1
done
