#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
for i in 1 2; do
echo "--- default"; python bench.py --workload train --steps 20 --warmup 3 --no-cpu-baseline --no-e2e | cut -c1-160
echo "--- no fused finalize"; WLSEG_NO_FUSED_FINALIZE=1 python bench.py --workload train --steps 20 --warmup 3 --no-cpu-baseline --no-e2e | cut -c1-160
echo "--- no zmask"; WLSEG_NO_ZMASK=1 python bench.py --workload train --steps 20 --warmup 3 --no-cpu-baseline --no-e2e | cut -c1-160
echo "--- neither"; WLSEG_NO_ZMASK=1 WLSEG_NO_FUSED_FINALIZE=1 python bench.py --workload train --steps 20 --warmup 3 --no-cpu-baseline --no-e2e | cut -c1-160
done
