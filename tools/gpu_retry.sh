#!/bin/bash
# gpurun with retries while the pod answers "transient" / busy (nothing is charged for those):
#   tools/gpu_retry.sh <timeout> [--gpus N] '<command>'
T=$1; shift
OPTS=()
if [ "$1" == "--gpus" ]; then OPTS=(--gpus "$2"); shift; shift; fi
for i in 1 2 3 4 5 6 7 8 9 10; do
    out=$(/usr/local/graft/bin/gpurun --timeout "$T" "${OPTS[@]}" -- "$@" 2>&1)
    echo "$out" | tail -150
    if echo "$out" | grep -q -E "status=transient|status=busy|rc=3"; then sleep 90; continue; fi
    break
done
