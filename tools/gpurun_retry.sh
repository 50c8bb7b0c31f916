#!/usr/bin/env bash
# usage: tools/gpurun_retry.sh <timeout-seconds> <command...>   — retries while the pod answers busy / transient
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun ${GPURUN_GPUS:+--gpus $GPURUN_GPUS} --timeout "$T" -- "$@" 2>&1)
  echo "$out"
  if echo "$out" | grep -q "status=transient\|status=busy\|rc=None"; then sleep 45; continue; fi
  break
done
