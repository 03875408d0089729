#!/bin/bash
# A/B of the medium / long horizon kernels (run on the GPU box): stock library, the runtime-horizon segmented
# kernels (OCD_RUNTIME_H=1, round 1's path) and the variant libraries of scripts/tuning/long_variants.txt.
cd "$(dirname "$0")/../.."
run() { echo -n "$1 "; shift; env "$@" python scripts/profile_point.py --H $H --C $C --B $B --reps 5; }
C=2
for pt in "15 1048576"; do set -- $pt; H=$1; B=$2
  run "H$H B$B runtime-seg wide" OCD_RUNTIME_H=1 OCD_KERNEL_FORM=wide
  run "H$H B$B stock(q104 cs sf1) wide" OCD_KERNEL_FORM=wide
  for v in q128sf1 q128sf2 q128sf3 q128sf5 q112sf2 q104sf2; do
    run "H$H B$B $v wide" OCD_B200_LIB=scratch/libocd_$v.so OCD_KERNEL_FORM=wide
  done
done
for pt in "50 1048576"; do set -- $pt; H=$1; B=$2
  run "H$H B$B runtime-seg wide" OCD_RUNTIME_H=1 OCD_KERNEL_FORM=wide
  run "H$H B$B stock(s5 r168 fr) wide" OCD_KERNEL_FORM=wide
  for v in s7r168 s6r168 s10r208 s10r168 s8r184; do
    run "H$H B$B $v wide" OCD_B200_LIB=scratch/libocd_$v.so OCD_KERNEL_FORM=wide
  done
done
C=6; H=15; B=1048576
run "H$H C6 B$B runtime-seg wide" OCD_RUNTIME_H=1 OCD_KERNEL_FORM=wide
run "H$H C6 B$B stock(q128 cs sf1) wide" OCD_KERNEL_FORM=wide
run "H$H C6 B$B stock(q128 cs sf1) auto"
for v in c6r104 c6r168; do run "H$H C6 B$B $v wide" OCD_B200_LIB=scratch/libocd_$v.so OCD_KERNEL_FORM=wide; done
C=4; run "H$H C4 B$B runtime-seg wide" OCD_RUNTIME_H=1 OCD_KERNEL_FORM=wide
run "H$H C4 B$B stock wide" OCD_KERNEL_FORM=wide
H=50; C=6; run "H$H C6 B$B runtime-seg wide" OCD_RUNTIME_H=1 OCD_KERNEL_FORM=wide
run "H$H C6 B$B stock wide" OCD_KERNEL_FORM=wide
