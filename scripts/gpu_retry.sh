#!/bin/bash
# gpurun with retries while the pod is busy (exit code 3 / "transient"): scripts/gpu_retry.sh <timeout> '<command>' [gpus]
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 40); do
  if [ "$G" = 1 ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$CMD" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$CMD" 2>&1); fi
  if echo "$OUT" | grep -q "status=transient\|nothing was charged — retry\|another call"; then sleep 45; continue; fi
  echo "$OUT"; exit 0
done
echo "gave up: pod busy"; exit 3
