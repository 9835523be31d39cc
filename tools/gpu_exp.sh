#!/bin/bash
# One B200 session: GPU test suite, then A/B of the step variants (same box, back to back).
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/tests_exp.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|Error" gpurun_out/tests_exp.log | tail -12
T="python bench.py --workload train --steps 20 --warmup 3 --no-cpu-baseline --no-e2e"
E="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e"
pick() { python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d['ms_per_step'],3), 'ms/step', round(d['value'],1), d['unit'], d['clocks'])" "$1" || echo "$1 FAILED"; }
WLSEG_NO_PDL=1 $T 2>gpurun_out/err_t0.log | pick "train no-pdl  "
$T 2>gpurun_out/err_t1.log | pick "train pdl     "
WLSEG_NO_PDL=1 $T 2>/dev/null | pick "train no-pdl 2"
$T 2>/dev/null | pick "train pdl 2   "
WLSEG_NO_PDL=1 WLSEG_POOL_GENERIC=1 $E 2>gpurun_out/err_e0.log | pick "eval no-pdl generic-pool"
WLSEG_NO_PDL=1 $E 2>/dev/null | pick "eval no-pdl   "
$E 2>gpurun_out/err_e1.log | pick "eval pdl      "
WLSEG_NO_PDL=1 $E 2>/dev/null | pick "eval no-pdl 2 "
$E 2>/dev/null | pick "eval pdl 2    "
