#!/bin/bash
# Round-2 state check: full GPU test suite, smoke, default bench line, eval + train timelines.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_default.json
python tools/step_timeline.py train > gpurun_out/timeline_train.txt 2>&1; head -12 gpurun_out/timeline_train.txt
python tools/step_timeline.py eval > gpurun_out/timeline_eval.txt 2>&1; head -14 gpurun_out/timeline_eval.txt
