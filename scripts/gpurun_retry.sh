#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3 = nothing charged).  Usage: scripts/gpurun_retry.sh <log> <timeout> <cmd...>
LOG=$1; TO=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
