#!/bin/bash
# usage: tools/gpu_retry.sh <log> <gpurun args...>   -- retries while the pod answers "busy" (exit code 3)
LOG=$1; shift
for i in $(seq 1 40); do
  gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$LOG"; then echo "done rc=$rc" >> "$LOG"; exit $rc; fi
  sleep 100
done
echo "gave up" >> "$LOG"
