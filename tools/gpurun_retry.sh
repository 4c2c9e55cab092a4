#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> <command string> [gpurun args...]
# retries a gpurun call while the pod answers "transient" (no box free; nothing charged)
T=$1; CMD=$2; shift 2
for i in $(seq 1 12); do
  OUT=$(/usr/local/graft/bin/gpurun "$@" --timeout "$T" -- "$CMD" 2>&1)
  if echo "$OUT" | grep -q "status=transient"; then
    sleep 90
    continue
  fi
  echo "$OUT"
  exit 0
done
echo "$OUT"
echo "gave up: still transient"
