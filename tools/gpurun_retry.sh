#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> [--gpus N] -- '<command>'   (retries while the pod answers "busy")
T=$1; shift
for i in $(seq 1 12); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout $T "$@" 2>&1)
  if echo "$OUT" | grep -q "status=transient"; then sleep 60; continue; fi
  echo "$OUT"; exit 0
done
echo "$OUT"; exit 3
