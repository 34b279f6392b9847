#!/bin/bash
# usage: tools/gpu_retry.sh <timeout> [--gpus N] -- '<command>' : retries while the pod answers busy (exit 3 / transient)
T=$1; shift
EXTRA=""
if [ "$1" = "--gpus" ]; then EXTRA="--gpus $2"; shift 2; fi
shift   # the "--"
for i in $(seq 1 40); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout $T $EXTRA -- "$@" 2>&1)
  echo "$OUT" | tail -60
  if echo "$OUT" | grep -q "status=transient\|retry in a few minutes\|no box"; then sleep 90; continue; fi
  break
done
