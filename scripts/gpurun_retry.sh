#!/bin/bash
# usage: scripts/gpurun_retry.sh <outfile> <gpurun args...>   -- retries while the pod has no free slot
out="$1"; shift
for i in $(seq 1 40); do
    gpurun "$@" > "$out" 2>&1
    rc=$?
    if grep -q "status=transient" "$out"; then sleep 90; continue; fi
    break
done
echo "gpurun_retry done rc=$rc attempts=$i" >> "$out"
