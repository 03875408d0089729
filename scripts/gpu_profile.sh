#!/bin/bash
# Run on the GPU box (via gpurun): launch list + full ncu capture of the solve kernel for the
# bench command.  Usage: scripts/gpu_profile.sh <tag> [bench args...]
set -u
TAG=${1:-r01}; shift || true
CMD="python bench.py --steps 5 --warmup 3 --no-extras $*"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_solve -s 4 -c 1 -f \
    -o gpurun_out/prof_solve_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
cut -c1-400 gpurun_out/plain_$TAG.log
