#!/bin/bash
# usage: tools/gpu_retry.sh <out-file> <gpurun args...>   -- retries while the pod answers "busy" (exit 3 / transient)
out="$1"; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  rc=$?
  if grep -q "status=transient" "$out" || [ $rc -eq 3 ]; then sleep 60; continue; fi
  exit $rc
done
exit 3
