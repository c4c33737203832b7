#!/bin/bash
# gpurun with retries while the pod answers "transient" (no slot free); usage: gpurun_retry.sh LOGFILE [gpurun args...]
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if ! grep -q "status=transient" "$log"; then exit 0; fi
  sleep 90
done
exit 3
