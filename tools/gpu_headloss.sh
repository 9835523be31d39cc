#!/bin/bash
mkdir -p gpurun_out
python tools/prof_headloss.py 5 2>&1 | grep -v Warn | tee gpurun_out/r2j_headloss.txt
ncu --set full --clock-control none --import-source on -k regex:"head_fwd|loss_fwd_bwd_kernel|loss_cols" --launch-skip 2 --launch-count 1 -o gpurun_out/r2j_head -f python tools/prof_headloss.py 1 > gpurun_out/ncu_head.log 2>&1; echo "ncu head rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"loss_fwd_bwd_kernel" --launch-skip 2 --launch-count 1 -o gpurun_out/r2j_loss -f python tools/prof_headloss.py 1 > gpurun_out/ncu_loss.log 2>&1; echo "ncu loss rc=$?"
ls -la gpurun_out/r2j*
