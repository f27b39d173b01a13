#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <gpus> '<command>'   -- retries while the pod answers "transient"/busy
T=$1; G=$2; shift 2
for i in $(seq 1 30); do
  if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" 2>&1); fi
  RC=$?
  if echo "$OUT" | grep -q "status=transient\|nothing was charged"; then sleep 90; continue; fi
  echo "$OUT" | tail -60
  exit $RC
done
echo "gave up after 30 tries"
