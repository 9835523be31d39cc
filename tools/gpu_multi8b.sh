#!/bin/bash
# usage: gpu_multi8b.sh N TAG  (under gpurun --gpus N): NCCL algorithm / protocol A/B on the training step, then the default
# bench line, BASELINE configs[3] (mixed 4 + 8 + 4, weak labels from lists) and configs[4] (Vistas 1080 x 1920) at N GPUs,
# and the in-situ timeline
N=${1:-8}; TAG=${2:-r2}
mkdir -p gpurun_out
pick() { python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d['ms_per_step'],3), 'ms/step', round(d['value'],1), d['unit'], 'e2e', d['e2e'] and round(d['e2e']['value'],1), '| eval', d.get('eval') and round(d['eval']['value'],1))" "$1" || echo "$1 FAILED"; }
tr() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 20 --warmup 3 --workload train --no-e2e --no-cpu-baseline --sustained-seconds 0 2>gpurun_out/${TAG}_ab_$1.err; }
tr 29521 > gpurun_out/${TAG}_n${N}_nccl_default.json; pick "nccl default" < gpurun_out/${TAG}_n${N}_nccl_default.json
NCCL_ALGO=NVLS tr 29522 > gpurun_out/${TAG}_n${N}_nccl_nvls.json; pick "NCCL_ALGO=NVLS" < gpurun_out/${TAG}_n${N}_nccl_nvls.json
NCCL_PROTO=Simple tr 29523 > gpurun_out/${TAG}_n${N}_nccl_simple.json; pick "NCCL_PROTO=Simple" < gpurun_out/${TAG}_n${N}_nccl_simple.json
NCCL_ALGO=Ring NCCL_PROTO=Simple NCCL_MAX_CTAS=8 tr 29524 > gpurun_out/${TAG}_n${N}_nccl_ring_simple_c8.json; pick "Ring Simple ctas8" < gpurun_out/${TAG}_n${N}_nccl_ring_simple_c8.json
grep -h "NCCL INFO.*Algo\|NVLS" gpurun_out/${TAG}_ab_*.err | head -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench N=$N rc=$?"; pick "default N=$N" < gpurun_out/${TAG}_bench_n$N.json; tail -2 gpurun_out/${TAG}_bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --workload train --mixed --boxes --no-cpu-baseline > gpurun_out/${TAG}_bench_mixed_n$N.json 2> gpurun_out/${TAG}_bench_mixed_n$N.err; echo "mixed rc=$?"; pick "mixed 4+8+4 N=$N" < gpurun_out/${TAG}_bench_mixed_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --workload train --dataset vistas --height 1080 --width 1920 --batch 2 --no-cpu-baseline > gpurun_out/${TAG}_bench_vistas_train_n$N.json 2> gpurun_out/${TAG}_bench_vistas_train_n$N.err; echo "vistas train rc=$?"; pick "vistas train N=$N" < gpurun_out/${TAG}_bench_vistas_train_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 --workload eval --dataset vistas --height 1080 --width 1920 --batch 4 --no-cpu-baseline > gpurun_out/${TAG}_bench_vistas_eval_n$N.json 2> gpurun_out/${TAG}_bench_vistas_eval_n$N.err; echo "vistas eval rc=$?"; pick "vistas eval N=$N" < gpurun_out/${TAG}_bench_vistas_eval_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 tools/step_timeline.py train > gpurun_out/${TAG}_timeline_train_n$N.txt 2>&1; echo "timeline rc=$?"; grep -v Warn gpurun_out/${TAG}_timeline_train_n$N.txt | grep "train:"; grep -A12 "NCCL kernels" gpurun_out/${TAG}_timeline_train_n$N.txt
