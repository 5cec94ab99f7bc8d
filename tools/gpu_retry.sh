#!/bin/bash
# usage: tools/gpu_retry.sh TIMEOUT_SECONDS 'command'   — gpurun with retries while the pod answers busy/transient
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|retry in a few minutes\|status=busy"; then
    sleep 60; continue
  fi
  echo "$out" | tail -40
  exit 0
done
echo "gpu_retry: gave up"
