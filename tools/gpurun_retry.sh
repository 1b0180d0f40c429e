#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3: nothing charged).
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>'
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
