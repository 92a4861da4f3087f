#!/bin/bash
# gpurun with retries while the pod reports "busy / draining" (nothing is charged for those answers).
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>' [gpus]
t=$1; cmd=$2; g=${3:-1}
for i in $(seq 1 40); do
  if [ "$g" = "1" ]; then out=$(/usr/local/graft/bin/gpurun --timeout $t -- "$cmd" 2>&1); else out=$(/usr/local/graft/bin/gpurun --gpus $g --timeout $t -- "$cmd" 2>&1); fi
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out" | tail -60; exit 0
done
echo "gave up after 40 transient answers"; exit 3
