#!/bin/bash
# usage: tools/gpushot.sh <log> <gpurun args...>   -- retries while the pod has no free GPU slot (exit 3), up to ~40 min
LOG=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1; rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$LOG"; then exit $rc; fi
  sleep 120
done
exit 3
