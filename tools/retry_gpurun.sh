#!/bin/bash
# usage: tools/retry_gpurun.sh <timeout_s> <command...>   -- retries while the pod answers "transient"
T=$1; shift
for i in $(seq 1 60); do
  out=$(gpurun --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy\|rc=3"; then sleep 150; continue; fi
  echo "$out" | tail -40; exit 0
done
echo "gave up"; exit 1
