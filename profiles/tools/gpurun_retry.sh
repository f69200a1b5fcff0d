#!/bin/bash
# usage: gpurun_retry.sh <timeout-seconds> '<command>'   -- retries while the pod answers "transient" (nothing charged)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 120; continue; fi
  echo "$out" | tail -80
  exit 0
done
echo "gave up after 40 transient answers"
