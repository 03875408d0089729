#!/bin/bash
# full ncu capture of the H = 15 (Q) and H = 50 (segmented) kernels -> gpurun_out/prof_r02_h{15,50}c{2,6}.ncu-rep
cd "$(dirname "$0")/../.."
for pt in ${POINTS:-"15 2 1048576" "50 2 262144" "15 6 262144"}; do set -- $pt
  CMD="python scripts/profile_point.py --H $1 --C $2 --B $3 --reps 2"
  $CMD > gpurun_out/point_r02_h$1c$2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_solve -s 1 -c 1 -f -o gpurun_out/prof_r02_h$1c$2 $CMD > gpurun_out/ncu_r02_h$1c$2.log 2>&1
  echo "H=$1 C=$2 rc=$? $(tail -1 gpurun_out/point_r02_h$1c$2.log)"
done
