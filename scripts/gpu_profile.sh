#!/bin/bash
# One GPU session's worth of profile artefacts for profiles/ (run on the GPU box via gpurun):
#   scripts/gpu_profile.sh <tag>      e.g. r02
# 1. bench.py --no-extras without ncu (must exit 0), 2. the launch list of the same command (gpu__time_duration),
# 3. ncu --set full of one k_solve launch of the bench (H = 5), and of the H = 15 / H = 50 / six-car sweep points,
# 4. the SASS of the bench kernel's hot loop.  Everything lands in gpurun_out/; scripts/refresh_profiles.py reduces it.
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r02}
O=gpurun_out
BENCH="python bench.py --steps 5 --warmup 3 --no-extras"
$BENCH > $O/bench_noextras_$TAG.json 2> $O/bench_noextras_$TAG.err || { echo "bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv $BENCH > $O/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_solve -s 3 -c 1 -f -o $O/prof_solve_$TAG $BENCH > $O/ncu_solve_$TAG.log 2>&1
echo "bench kernel rc=$?"
for pt in "15 2 1048576 h15c2" "50 2 1048576 h50c2" "5 6 1048576 h5c6" "15 6 1048576 h15c6"; do set -- $pt
  CMD="python scripts/profile_point.py --H $1 --C $2 --B $3 --reps 2"
  $CMD > $O/point_${TAG}_$4.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_solve -s 1 -c 1 -f -o $O/prof_${TAG}_$4 $CMD > $O/ncu_${TAG}_$4.log 2>&1
  echo "$4 rc=$? $(tail -1 $O/point_${TAG}_$4.log)"
done
