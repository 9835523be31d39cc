#!/bin/bash
mkdir -p gpurun_out
python tools/prof_headloss.py 5 2>&1 | grep -v Warn | tee gpurun_out/hl_default.txt
WLSEG_LOSS_COLS=all python tools/prof_headloss.py 5 2>&1 | grep -v Warn | tee gpurun_out/hl_cols_all.txt
ncu --set full --clock-control none --import-source on -k regex:"loss_fwd_bwd_kernel" --launch-skip 2 --launch-count 1 -o gpurun_out/hl_loss_tile -f python tools/prof_headloss.py 1 > gpurun_out/ncu_loss.log 2>&1; echo "ncu loss tile rc=$?"
WLSEG_LOSS_COLS=all ncu --set full --clock-control none --import-source on -k regex:"loss_cols" --launch-skip 2 --launch-count 1 -o gpurun_out/hl_loss_cols -f python tools/prof_headloss.py 1 > gpurun_out/ncu_loss2.log 2>&1; echo "ncu loss cols rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"head_fwd" --launch-skip 2 --launch-count 1 -o gpurun_out/hl_head -f python tools/prof_headloss.py 1 > gpurun_out/ncu_head.log 2>&1; echo "ncu head rc=$?"
ls -la gpurun_out/hl_*
