#!/bin/bash
# One B200 session: conv / network tests, then A/B of the epilogue schedules (same box, back to back).
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_network.py tests/test_gpu_train.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/tests_exp.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_exp.log
T="python bench.py --workload train --steps 20 --warmup 3 --no-cpu-baseline --no-e2e"
E="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e"
pick() { python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d['ms_per_step'],3), 'ms/step', round(d['value'],1), d['unit'], d['clocks']['sm_mhz'])" "$1" || echo "$1 FAILED"; }
WLSEG_RES_MID=0 WLSEG_EPI_DB=0 $E 2>/dev/null | pick "eval  mid=0 db=0"
WLSEG_RES_MID=1 WLSEG_EPI_DB=0 $E 2>/dev/null | pick "eval  mid=1 db=0"
WLSEG_RES_MID=0 WLSEG_EPI_DB=1 $E 2>/dev/null | pick "eval  mid=0 db=1"
WLSEG_RES_MID=1 WLSEG_EPI_DB=1 $E 2>/dev/null | pick "eval  mid=1 db=1"
WLSEG_RES_MID=0 WLSEG_EPI_DB=0 $E 2>/dev/null | pick "eval  mid=0 db=0 (2)"
WLSEG_RES_MID=1 WLSEG_EPI_DB=1 $E 2>/dev/null | pick "eval  mid=1 db=1 (2)"
WLSEG_RES_MID=0 WLSEG_EPI_DB=0 $T 2>/dev/null | pick "train mid=0 db=0"
WLSEG_RES_MID=1 WLSEG_EPI_DB=1 $T 2>/dev/null | pick "train mid=1 db=1"
WLSEG_RES_MID=0 WLSEG_EPI_DB=0 python tools/layer_table.py eval > gpurun_out/layers_eval_00.txt 2>&1
WLSEG_RES_MID=1 WLSEG_EPI_DB=1 python tools/layer_table.py eval > gpurun_out/layers_eval_11.txt 2>&1
