#!/bin/bash
# Runs every GPU test file in its own process (a trapped kernel poisons its CUDA context, not the
# others) under a timeout and collects logs in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
rc_all=0
for f in ${@:-tests/test_gpu_confmat.py tests/test_gpu_head_loss.py tests/test_gpu_misc.py tests/test_gpu_conv.py tests/test_gpu_network.py}; do
  name=$(basename $f .py)
  timeout 600 python -m pytest $f -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/$name.log 2>&1
  rc=$?
  echo "== $f rc=$rc"; tail -25 gpurun_out/$name.log
  [ $rc -ne 0 ] && rc_all=1
done
exit $rc_all
