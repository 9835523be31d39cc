#!/bin/bash
# usage: gpu_multi8.sh N TAG  (under gpurun --gpus N): A/B of the SM partition between convolutions and NCCL, then the
# default bench line and the in-situ timeline with the chosen defaults
N=${1:-8}; TAG=${2:-r2}
mkdir -p gpurun_out
pick() { python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d['ms_per_step'],3), 'ms/step', round(d['value'],1), d['unit'], 'e2e', d['e2e'] and round(d['e2e']['value'],1), '| eval', d.get('eval') and round(d['eval']['value'],1))" "$1" || echo "$1 FAILED"; }
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 20 --warmup 3 --workload train --no-e2e --no-cpu-baseline --sustained-seconds 0 2>/dev/null; }
WLSEG_CONV_SMS=148 run 29521 > gpurun_out/${TAG}_n${N}_sms148.json; pick "sms148 nccl-default" < gpurun_out/${TAG}_n${N}_sms148.json
WLSEG_CONV_SMS=144 NCCL_MAX_CTAS=4 run 29522 > gpurun_out/${TAG}_n${N}_sms144_c4.json; pick "sms144 ctas4" < gpurun_out/${TAG}_n${N}_sms144_c4.json
WLSEG_CONV_SMS=140 NCCL_MAX_CTAS=8 run 29523 > gpurun_out/${TAG}_n${N}_sms140_c8.json; pick "sms140 ctas8" < gpurun_out/${TAG}_n${N}_sms140_c8.json
WLSEG_CONV_SMS=144 run 29524 > gpurun_out/${TAG}_n${N}_sms144.json; pick "sms144 nccl-default" < gpurun_out/${TAG}_n${N}_sms144.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench N=$N rc=$?"; pick "default N=$N" < gpurun_out/${TAG}_bench_n$N.json; tail -2 gpurun_out/${TAG}_bench_n$N.err
NCCL_MAX_CTAS=4 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/step_timeline.py train > gpurun_out/${TAG}_timeline_train_n$N.txt 2>&1; echo "timeline rc=$?"; grep -v Warn gpurun_out/${TAG}_timeline_train_n$N.txt | grep "train:"; grep -A12 "NCCL kernels" gpurun_out/${TAG}_timeline_train_n$N.txt
python bench.py --workload train --no-e2e --no-cpu-baseline --sustained-seconds 0 > gpurun_out/${TAG}_n1_on_n${N}box.json 2>/dev/null; pick "N=1 on this box" < gpurun_out/${TAG}_n1_on_n${N}box.json
