#!/bin/bash
# fused BN-backward reduction: parity tests, then A/B of the training step on one box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_train.py tests/test_gpu_baseline_shapes.py -m gpu -x -q -s -k "bnbwd or fprop_bn or fused_bn_backward or train_step or graph_replay" > gpurun_out/bnb_tests.log 2>&1; echo "tests rc=$?"; grep -E "fused vs|passed|failed|Error|error" gpurun_out/bnb_tests.log | tail -20
for f in 0 1; do
  WLSEG_BNB_FUSE=$f python bench.py --workload train --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --sustained-seconds 0 > gpurun_out/bnb_bench_$f.json 2> gpurun_out/bnb_bench_$f.err; echo "fuse=$f rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bnb_bench_$f.json')); print(d['value'], d['ms_per_step'])"
done
for f in 0 1; do
  WLSEG_BNB_FUSE=$f python bench.py --workload train --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --sustained-seconds 0 > gpurun_out/bnb_bench_${f}b.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/bnb_bench_${f}b.json')); print('rerun fuse=$f', d['value'], d['ms_per_step'])"
done
python tools/step_timeline.py train > gpurun_out/bnb_timeline_train.txt 2>&1; head -24 gpurun_out/bnb_timeline_train.txt | tail -22
for f in 0 1; do
  WLSEG_BN_FIN_IN_CONV=$f python bench.py --workload train --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --sustained-seconds 0 > gpurun_out/fin_bench_$f.json 2> gpurun_out/fin_bench_$f.err; python -c "
import json; d=json.load(open('gpurun_out/fin_bench_$f.json')); print('fin_in_conv=$f', d['value'], d['ms_per_step'])"
done
