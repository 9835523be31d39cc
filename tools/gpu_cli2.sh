#!/bin/bash
# CLI trio on one GPU with this round's options: checkpoints (EMA restore), void replacement, system-size outputs,
# PNG exports, and the optional model parts (PSP + hybrid upsampler + group norm) through train -> evaluate.
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; exit 1; }
PD=iv2019-boosting-semantic-segmentation-with-weak-labels_b200/wlseg/problem_definitions/cityscapes/problem01.json
S="--synthetic --height_feature_extractor 256 --width_feature_extractor 512"
rm -rf /tmp/wl1 /tmp/wl4 /tmp/wlpred /tmp/wlout; mkdir -p /tmp/wlpred /tmp/wlout
python train.py /tmp/wl1 cityscapes $S --steps 6 --save_checkpoints_steps 3 > gpurun_out/cli2_train.log 2>&1; echo "train rc=$?"; tail -2 gpurun_out/cli2_train.log; ls /tmp/wl1 | head
python evaluate.py /tmp/wl1 8 $PD synthetic cityscapes $S --Nb 2 --restore_emas --replace_voids > gpurun_out/cli2_eval.log 2>&1; echo "evaluate (EMA, replace_voids) rc=$?"; grep -E "mIoU|mean|accuracy" gpurun_out/cli2_eval.log | tail -3
python evaluate.py /tmp/wl1 8 $PD synthetic cityscapes $S --Nb 2 --eval_all_ckpts > gpurun_out/cli2_eval_all.log 2>&1; echo "evaluate --eval_all_ckpts rc=$?"; grep -c "checkpoint" gpurun_out/cli2_eval_all.log
python predict.py /tmp/wl1 $PD /tmp/wlpred cityscapes $S --height_system 300 --width_system 500 --replace_voids --export_color_decisions --export_lids_images --results_dir /tmp/wlout > gpurun_out/cli2_predict.log 2>&1; echo "predict rc=$?"; tail -2 gpurun_out/cli2_predict.log; ls /tmp/wlout | head -4; python -c "
from PIL import Image; import glob; f=sorted(glob.glob('/tmp/wlout/*color.png'))[0]; print(f, Image.open(f).size)"
python train.py /tmp/wl4 cityscapes $S --steps 4 --save_checkpoints_steps 4 --psp_module --upsampling_method hybrid --norm_layer group --fov_expansion_kernel_size 3 --fov_expansion_kernel_rate 2 > gpurun_out/cli2_train_opts.log 2>&1; echo "train (psp+hybrid+group+fov) rc=$?"; tail -2 gpurun_out/cli2_train_opts.log
python evaluate.py /tmp/wl4 4 $PD synthetic cityscapes $S --Nb 2 --psp_module --upsampling_method hybrid --norm_layer group --fov_expansion_kernel_size 3 --fov_expansion_kernel_rate 2 > gpurun_out/cli2_eval_opts.log 2>&1; echo "evaluate (psp+hybrid+group+fov) rc=$?"; tail -3 gpurun_out/cli2_eval_opts.log
