#!/bin/bash
mkdir -p gpurun_out
for m in 0 1; do echo "--- WLSEG_PAIR=$m"; WLSEG_PAIR=$m python tools/prof_conv.py 5 pairprobe; done 2>&1 | grep -v Warn | tee gpurun_out/r2c_pairprobe.txt
WLSEG_PAIR=1 ncu --set full --clock-control none --import-source on -k regex:"conv_igemm" --launch-skip 1 --launch-count 1 -o gpurun_out/r2c_pair_256_1024res -f python tools/prof_conv.py 2 pairprobe > gpurun_out/ncu_pair.log 2>&1; echo "ncu pair rc=$?"
WLSEG_PAIR=0 ncu --set full --clock-control none --import-source on -k regex:"conv_igemm" --launch-skip 1 --launch-count 1 -o gpurun_out/r2c_single_256_1024res -f python tools/prof_conv.py 2 pairprobe > gpurun_out/ncu_single.log 2>&1; echo "ncu single rc=$?"
ls -la gpurun_out/*.ncu-rep
