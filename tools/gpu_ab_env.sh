#!/bin/bash
# usage: gpu_ab_env.sh VAR A B  -> eval and train step with VAR=A and VAR=B, twice, on one box
V=$1; A=$2; B=$3
for x in $A $B $A $B; do
  env $V=$x python bench.py --workload eval --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --sustained-seconds 0 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); print('eval  $V=$x', round(d['ms_per_step'],3), 'ms/step', round(d['value'],1), d['unit'])"
  env $V=$x python bench.py --workload train --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --sustained-seconds 0 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); print('train $V=$x', round(d['ms_per_step'],3), 'ms/step', round(d['value'],1), d['unit'])"
done
