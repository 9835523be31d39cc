#!/bin/bash
# One B200 session: BN / train tests, then A/B of the traversal direction of the BN passes (same box, back to back).
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_misc.py tests/test_gpu_train.py tests/test_gpu_psp.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/tests_exp.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_exp.log
T="python bench.py --workload train --steps 20 --warmup 3 --no-cpu-baseline --no-e2e"
pick() { python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d['ms_per_step'],3), 'ms/step', round(d['value'],1), d['unit'], d['clocks']['sm_mhz'])" "$1" || echo "$1 FAILED"; }
WLSEG_BN_RED_REV=0 WLSEG_BN_APPLY_REV=0 $T 2>/dev/null | pick "train red=asc  apply=asc "
WLSEG_BN_RED_REV=1 WLSEG_BN_APPLY_REV=0 $T 2>/dev/null | pick "train red=desc apply=asc "
WLSEG_BN_RED_REV=0 WLSEG_BN_APPLY_REV=1 $T 2>/dev/null | pick "train red=asc  apply=desc"
WLSEG_BN_RED_REV=1 WLSEG_BN_APPLY_REV=1 $T 2>/dev/null | pick "train red=desc apply=desc"
WLSEG_BN_RED_REV=0 WLSEG_BN_APPLY_REV=0 $T 2>/dev/null | pick "train red=asc  apply=asc  (2)"
WLSEG_BN_RED_REV=1 WLSEG_BN_APPLY_REV=1 $T 2>/dev/null | pick "train red=desc apply=desc (2)"
