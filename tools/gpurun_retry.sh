#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3).  Usage: tools/gpurun_retry.sh LOGFILE [gpurun args...]
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
