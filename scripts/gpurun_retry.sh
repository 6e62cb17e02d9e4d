#!/bin/bash
# usage: gpurun_retry.sh <gpus> <timeout> <command...>   -- retries while the pod answers "busy" (rc 3 / transient)
G=$1; shift; TO=$1; shift
for i in $(seq 1 30); do
  if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout $TO -- "$@" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $TO -- "$@" 2>&1); fi
  echo "$OUT" | tail -60
  if echo "$OUT" | grep -q "status=transient\|nothing was charged"; then sleep 120; continue; fi
  break
done
