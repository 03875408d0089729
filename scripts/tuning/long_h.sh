#!/bin/bash
# A/B of the medium / long horizon kernels (run on the GPU box): stock library, the runtime-horizon segmented
# kernels (OCD_RUNTIME_H=1, round 1's path) and the variant libraries of scripts/tuning/long_variants.txt.
cd "$(dirname "$0")/../.."
run() { echo -n "$1 "; shift; env "$@" python scripts/profile_point.py --H $H --C 2 --B $B --reps 5; }
for pt in "15 262144" "15 1048576"; do set -- $pt; H=$1; B=$2
  run "H$H B$B runtime-seg wide" OCD_RUNTIME_H=1 OCD_KERNEL_FORM=wide
  run "H$H B$B stock(q128) wide" OCD_KERNEL_FORM=wide
  run "H$H B$B stock(q128) latency" OCD_KERNEL_FORM=latency
  for v in q168 q168sf q128sf q128cs q128cssf q104cssf q96cssf; do
    run "H$H B$B $v wide" OCD_B200_LIB=scratch/libocd_$v.so OCD_KERNEL_FORM=wide
  done
  run "H$H B$B q168sf latency" OCD_B200_LIB=scratch/libocd_q168sf.so OCD_KERNEL_FORM=latency
  run "H$H B$B q128cssf throughput" OCD_B200_LIB=scratch/libocd_q128cssf.so OCD_KERNEL_FORM=throughput
done
for pt in "50 65536" "50 262144"; do set -- $pt; H=$1; B=$2
  run "H$H B$B runtime-seg wide" OCD_RUNTIME_H=1 OCD_KERNEL_FORM=wide
  run "H$H B$B stock(s168 fr) wide" OCD_KERNEL_FORM=wide
  run "H$H B$B stock(s168 fr) latency" OCD_KERNEL_FORM=latency
  run "H$H B$B stock(s168 fr) throughput" OCD_KERNEL_FORM=throughput
  for v in s128 s168nofr s128nofr; do
    run "H$H B$B $v wide" OCD_B200_LIB=scratch/libocd_$v.so OCD_KERNEL_FORM=wide
  done
done
