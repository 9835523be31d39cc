#!/bin/bash
# usage: gpu_multi.sh N   (run under gpurun --gpus N): the 2-rank GPU tests (N >= 2), then the eval and train benches
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
if [ "$N" = "2" ]; then
  timeout 400 python -m pytest tests/test_gpu_cross_replica.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests_n2.log 2>&1; echo "2-GPU tests rc=$?"; tail -3 gpurun_out/tests_n2.log
fi
pick() { python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d['ms_per_step'],3), 'ms/step', round(d['value'],1), d['unit'], 'e2e', d['e2e'] and round(d['e2e']['value'],1))" "$1" || echo "$1 FAILED"; }
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_eval_n$N.json 2> gpurun_out/bench_eval_n$N.err; echo "eval N=$N rc=$?"; pick "eval N=$N" < gpurun_out/bench_eval_n$N.json; tail -3 gpurun_out/bench_eval_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload train --steps 10 --warmup 3 > gpurun_out/bench_train_n$N.json 2> gpurun_out/bench_train_n$N.err; echo "train N=$N rc=$?"; pick "train N=$N" < gpurun_out/bench_train_n$N.json; tail -3 gpurun_out/bench_train_n$N.err
