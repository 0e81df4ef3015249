#!/bin/bash
# usage: tools/gpu/submit.sh <log> <gpurun args...>   -- retries while the pod has no free slot (gpurun exit code 3)
LOG=$1; shift
for attempt in $(seq 1 12); do
  gpurun "$@" > "$LOG" 2>&1; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 75
done
exit 3
