#!/bin/bash
# gpurun with retries while the pod answers "transient" (nothing is charged for those): tools/gpu_retry.sh <timeout> '<command>'
T=$1; shift
for i in 1 2 3 4 5 6 7 8; do
    out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
    echo "$out" | tail -150
    if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
    break
done
