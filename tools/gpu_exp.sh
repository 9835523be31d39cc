#!/bin/bash
# One B200 session: conv / network tests, then A/B of the alternating tile order of the inference convolutions.
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_network.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/tests_exp.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_exp.log
E="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e"
pick() { python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d['ms_per_step'],3), 'ms/step', round(d['value'],1), d['unit'], d['clocks']['sm_mhz'])" "$1" || echo "$1 FAILED"; }
WLSEG_SNAKE=0 $E 2>/dev/null | pick "eval same direction "
WLSEG_SNAKE=1 $E 2>/dev/null | pick "eval alternating    "
WLSEG_SNAKE=0 $E 2>/dev/null | pick "eval same direction 2"
WLSEG_SNAKE=1 $E 2>/dev/null | pick "eval alternating    2"
