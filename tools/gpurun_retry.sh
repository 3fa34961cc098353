#!/bin/bash
# gpurun with retries while the pod has no free slot (exit 3 / "transient"): usage tools/gpurun_retry.sh <timeout> [--gpus N] -- <cmd>
tmo=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  out=$(/usr/local/graft/bin/gpurun --timeout $tmo "$@" 2>&1)
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient"; then sleep 120; continue; fi
  break
done
