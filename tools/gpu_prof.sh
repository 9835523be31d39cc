#!/bin/bash
# per-layer tables (no profiler) + ncu --set full of representative conv launches
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
python tools/layer_table.py eval > gpurun_out/layers_eval.txt 2>&1; echo "layers eval rc=$?"
python tools/layer_table.py train > gpurun_out/layers_train.txt 2>&1; echo "layers train rc=$?"
python tools/prof_conv.py 3 > gpurun_out/prof_conv_plain.txt 2>&1; echo "prof_conv rc=$?"; cat gpurun_out/prof_conv_plain.txt
ncu --set full --clock-control none --import-source on -k regex:"conv_igemm|conv_wgrad" --launch-skip 1 --launch-count 16 -o gpurun_out/conv_full -f python tools/prof_conv.py 2 > gpurun_out/ncu_conv_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/ncu_conv_full.log
