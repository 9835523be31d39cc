#!/bin/bash
# halo form of the narrow convolutions: A/B of the evaluation and the training step on one box, then the affected tests
mkdir -p gpurun_out
for h in 0 1 0 1; do
  WLSEG_HALO=$h python bench.py --workload eval --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --sustained-seconds 0 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); print('eval  halo=$h', round(d['ms_per_step'],3), 'ms/step', round(d['value'],1), d['unit'])"
  WLSEG_HALO=$h python bench.py --workload train --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --sustained-seconds 0 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); print('train halo=$h', round(d['ms_per_step'],3), 'ms/step', round(d['value'],1), d['unit'])"
done
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_network.py tests/test_gpu_baseline_shapes.py -m gpu -x -q 2>&1 | tail -3
