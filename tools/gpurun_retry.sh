#!/bin/bash
# gpurun with retries while the pod answers "transient" (busy): tools/gpurun_retry.sh <timeout> '<command>' [extra gpurun flags]
T=$1; CMD=$2; shift 2
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun "$@" --timeout "$T" -- "$CMD" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
