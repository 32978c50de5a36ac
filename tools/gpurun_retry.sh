#!/bin/bash
# gpurun with retries while the pod is busy (exit code 3 = nothing charged):  tools/gpurun_retry.sh <gpurun args ...>
for attempt in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun "$@"; rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 150
done
exit 3
