#!/bin/bash
# usage: tools/gpu.sh <logfile> <timeout-seconds> <command...>   - gpurun with retries while the pod is busy
log=$1; shift; to=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient\|rc=3\|no box\|busy" "$log" && [ $rc -ne 0 ] && ! grep -q "charged=[1-9]" "$log"; then
    sleep 90; continue
  fi
  break
done
echo "gpurun rc=$rc" >> "$log"
