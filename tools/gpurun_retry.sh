#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3):  tools/gpurun_retry.sh <log> <gpurun args...>
log=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
