#!/bin/bash
# usage: gpu_multi2.sh N TAG  (under gpurun --gpus N): 2-rank GPU tests (N == 2), the default bench line at N (training
# headline with the gradient all-reduce in the timed region, evaluation nested), an in-situ timeline with the NCCL kernels
N=${1:-2}; TAG=${2:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_cross_replica.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/${TAG}_tests_n2.log 2>&1; echo "2-GPU tests rc=$?"; tail -3 gpurun_out/${TAG}_tests_n2.log
fi
pick() { python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d['ms_per_step'],3), 'ms/step', round(d['value'],1), d['unit'], 'e2e', d['e2e'] and round(d['e2e']['value'],1), '| eval', d.get('eval') and round(d['eval']['value'],1))" "$1" || echo "$1 FAILED"; }
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench N=$N rc=$?"; pick "N=$N" < gpurun_out/${TAG}_bench_n$N.json; tail -3 gpurun_out/${TAG}_bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/step_timeline.py train > gpurun_out/${TAG}_timeline_train_n$N.txt 2>&1; echo "timeline rc=$?"; grep -v Warn gpurun_out/${TAG}_timeline_train_n$N.txt | head -14; grep -A40 "NCCL kernels" gpurun_out/${TAG}_timeline_train_n$N.txt
if [ -n "$3" ]; then
  for v in $3; do
    echo "--- NCCL_MAX_CTAS=$v"; NCCL_MAX_CTAS=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 20 --warmup 3 --workload train --no-e2e --no-cpu-baseline --sustained-seconds 0 > gpurun_out/${TAG}_bench_n${N}_ctas$v.json 2>/dev/null; pick "N=$N ctas=$v" < gpurun_out/${TAG}_bench_n${N}_ctas$v.json
  done
fi
