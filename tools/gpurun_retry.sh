#!/bin/bash
# Retry a gpurun call while the pod answers "busy" (exit 3: nothing charged).  usage: tools/gpurun_retry.sh <log> <gpurun args...>
LOG=$1; shift
for i in $(seq 1 40); do
  timeout 3500 gpurun "$@" > "$LOG" 2>&1
  rc=$?
  echo "exit $rc (attempt $i)" >> "$LOG"
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
