#!/bin/bash
# usage: tools/gpurun_retry_n.sh <gpus> <timeout> <command...>
G=$1; T=$2; shift; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@" > /tmp/gpurun_try_n.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_try_n.log; then sleep 120; continue; fi
  cat /tmp/gpurun_try_n.log
  exit $rc
done
cat /tmp/gpurun_try_n.log
exit 3
