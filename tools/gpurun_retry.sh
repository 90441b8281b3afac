#!/bin/bash
# gpurun with retries while the pod answers "transient" (no GPU slot free, nothing charged)
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>'
T=$1; shift
for i in $(seq 1 20); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  if echo "$OUT" | grep -q 'status=transient'; then sleep 150; continue; fi
  echo "$OUT"; exit 0
done
echo "$OUT"; exit 3
