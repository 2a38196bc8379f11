#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> <command...>   - retries gpurun while the pod answers busy (exit 3 / transient)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  rc=$?
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then sleep 60; continue; fi
  echo "$out"
  exit $rc
done
echo "gave up"; exit 3
