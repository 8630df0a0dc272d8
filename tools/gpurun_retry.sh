#!/bin/bash
# usage: gpurun_retry.sh <log> <gpurun args...>   retries while the pod answers "transient" (nothing charged)
LOG=$1; shift
for attempt in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  if ! grep -q "status=transient" "$LOG"; then break; fi
  sleep 150
done
