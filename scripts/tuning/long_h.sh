#!/bin/bash
# A/B of the medium / long horizon kernels (run on the GPU box): stock library against the variant libraries of
# scripts/tuning/long_variants.txt (built by build_variants.sh into scratch/).
cd "$(dirname "$0")/../.."
run() { echo -n "$1 "; shift; env "$@" python scripts/profile_point.py --H $H --C $C --B $B --reps 4; }
C=2; H=15; B=1048576
run "H$H stock(q104 sf1)" OCD_KERNEL_FORM=wide
for v in q128sf1 q128sf2 q128sf3 q128sf5 q112sf2 q104sf2 q96sf1 q88sf1 segc15 segc15r128; do
  run "H$H $v" OCD_B200_LIB=scratch/libocd_$v.so OCD_KERNEL_FORM=wide
done
H=50
run "H$H stock(s5 r168)" OCD_KERNEL_FORM=wide
for v in s7r168 s5r128 s5r152 s10r208; do
  run "H$H $v" OCD_B200_LIB=scratch/libocd_$v.so OCD_KERNEL_FORM=wide
done
C=3; H=15
run "H$H C3 stock(segc)" OCD_KERNEL_FORM=wide
run "H$H C3 q2c3 (Q kernel)" OCD_B200_LIB=scratch/libocd_q2c3.so OCD_KERNEL_FORM=wide
