#!/bin/bash
# gpurun with retries while the pod answers "transient" / busy (nothing is charged for those).
# usage: tools/grun.sh <timeout-seconds> [--gpus N] -- '<command>'
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T "$@" > /tmp/grun.$$.log 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" /tmp/grun.$$.log || [ $rc -eq 3 ]; then
    sleep 60
    continue
  fi
  cat /tmp/grun.$$.log
  exit $rc
done
cat /tmp/grun.$$.log
exit 3
