#!/bin/bash
# usage: gpu_multi.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_eval_n$N.json 2> gpurun_out/bench_eval_n$N.err; echo "eval N=$N rc=$?"; cat gpurun_out/bench_eval_n$N.json; tail -5 gpurun_out/bench_eval_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload train --steps 10 --warmup 3 > gpurun_out/bench_train_n$N.json 2> gpurun_out/bench_train_n$N.err; echo "train N=$N rc=$?"; cat gpurun_out/bench_train_n$N.json; tail -5 gpurun_out/bench_train_n$N.err
