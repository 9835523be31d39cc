#!/bin/bash
# The reference-facing CLI trio on synthetic inputs: 1 GPU, then (if available) 2 GPUs incl. --cross_replica_norm.
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; exit 1; }
PD=iv2019-boosting-semantic-segmentation-with-weak-labels_b200/wlseg/problem_definitions/cityscapes/problem01.json
rm -rf /tmp/wl1 /tmp/wl2 /tmp/wl3 /tmp/wlpred; mkdir -p /tmp/wlpred
python train.py /tmp/wl1 cityscapes --synthetic --steps 6 --height_feature_extractor 256 --width_feature_extractor 512 > gpurun_out/cli_train.log 2>&1; echo "train rc=$?"; tail -4 gpurun_out/cli_train.log
python evaluate.py /tmp/wl1 8 $PD synthetic cityscapes --synthetic --Nb 2 --height_feature_extractor 256 --width_feature_extractor 512 > gpurun_out/cli_eval.log 2>&1; echo "evaluate rc=$?"; tail -4 gpurun_out/cli_eval.log
python predict.py /tmp/wl1 $PD /tmp/wlpred cityscapes --synthetic --height_feature_extractor 256 --width_feature_extractor 512 > gpurun_out/cli_predict.log 2>&1; echo "predict rc=$?"; tail -3 gpurun_out/cli_predict.log
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
timeout 200 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 train.py /tmp/wl2 cityscapes --synthetic --distribute --steps 6 --height_feature_extractor 256 --width_feature_extractor 512 > gpurun_out/cli_train2.log 2>&1; echo "train x2 rc=$?"; tail -4 gpurun_out/cli_train2.log
timeout 200 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 train.py /tmp/wl3 cityscapes --synthetic --distribute --cross_replica_norm --steps 6 --height_feature_extractor 256 --width_feature_extractor 512 > gpurun_out/cli_train2x.log 2>&1; echo "train x2 sync-BN rc=$?"; tail -4 gpurun_out/cli_train2x.log
python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 evaluate.py /tmp/wl2 8 $PD synthetic cityscapes --synthetic --Nb 2 --height_feature_extractor 256 --width_feature_extractor 512 > gpurun_out/cli_eval2.log 2>&1; echo "evaluate x2 rc=$?"; tail -4 gpurun_out/cli_eval2.log
fi
