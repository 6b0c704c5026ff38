#!/bin/bash
# usage: gpurun_retry.sh <timeout> <command...>: retries while the pod answers "transient" (nothing charged)
T=$1; shift
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|rc=3"; then sleep 150; continue; fi
  echo "$out"; exit 0
done
echo "$out"; echo "gave up after retries"
