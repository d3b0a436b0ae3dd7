#!/bin/bash
# gpurun with retries while the pod answers "transient" (no slot / no box: nothing is charged).  usage: gpurun_retry.sh <log> <gpurun args...>
log=$1; shift
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if grep -q "status=transient" "$log" || grep -q "rc=3" "$log"; then sleep 90; continue; fi
  break
done
