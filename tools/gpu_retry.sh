#!/bin/bash
# usage: [GPUS=N] tools/gpu_retry.sh TIMEOUT_SECONDS 'command'   — gpurun with retries while the pod answers busy/transient
T=$1; shift
G=""
if [ -n "$GPUS" ]; then G="--gpus $GPUS"; fi
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun $G --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|retry in a few minutes\|status=busy\|no box or slot"; then
    sleep 60; continue
  fi
  echo "$out" | tail -60
  exit 0
done
echo "gpu_retry: gave up"
