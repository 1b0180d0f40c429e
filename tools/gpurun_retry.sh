#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <command...>   -- retries while the pod answers busy (rc 3), nothing is charged for those
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > /tmp/gpurun_try.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_try.log; then sleep 90; continue; fi
  cat /tmp/gpurun_try.log
  exit $rc
done
cat /tmp/gpurun_try.log
exit 3
