#!/bin/bash
# gpurun with retries while the pod answers "transient" (no slot free; nothing charged).
#   tools/gpurun_retry.sh <log> <gpurun args...>
LOG=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  if ! grep -q "status=transient" "$LOG"; then break; fi
  sleep 90
done
