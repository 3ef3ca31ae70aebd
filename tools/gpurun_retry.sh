#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> '<command>' [extra gpurun flags]   -- retries while the pod answers busy/transient (nothing is charged then)
T=$1; shift; CMD=$1; shift
for i in $(seq 1 40); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout $T "$@" -- "$CMD" 2>&1)
  if echo "$OUT" | grep -q "status=transient\|exit code 3\|status=busy\|rc=3"; then sleep 60; continue; fi
  echo "$OUT"; exit 0
done
echo "$OUT"; exit 3
