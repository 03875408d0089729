#!/bin/bash
# Run on the GPU box (via gpurun): full ncu capture of the solve kernel at chosen sweep points.
# Usage: scripts/gpu_profile_sweep.sh <tag> "H C B" ["H C B" ...]
set -u
TAG=${1:-r01}; shift
mkdir -p gpurun_out
for pt in "$@"; do
    set -- $pt; H=$1; C=$2; B=$3
    CMD="python scripts/profile_point.py --H $H --C $C --B $B"
    $CMD > gpurun_out/point_${TAG}_h${H}c${C}.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:k_solve -s 1 -c 1 -f \
        -o gpurun_out/prof_${TAG}_h${H}c${C} $CMD > gpurun_out/ncu_${TAG}_h${H}c${C}.log 2>&1
    echo "H=$H C=$C rc=$? $(cat gpurun_out/point_${TAG}_h${H}c${C}.log | tail -1)"
done
