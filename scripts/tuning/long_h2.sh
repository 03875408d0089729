#!/bin/bash
# Second A/B of the H = 15 Q kernel (GPU box): more registers for more interleaving of the unrolled steps.
# The forward sweep's stall samples are fixed-latency waits and MUFU / shared-memory scoreboards (ILP), not
# instruction fetch, so: fences every 3 / 5 / no steps at 136 registers (still five blocks per SM) and 168 (four).
# Variants: scripts/tuning/build_variants.sh <name> solve_15_1_3 "-DOCD_Q_REGS_ALL=<r> -DOCD_Q_SF=<n>"
cd "$(dirname "$0")/../.."
run() { echo -n "$1 "; shift; env "$@" python scripts/profile_point.py --H $H --C $C --B $B --reps 4; }
C=2; H=15; B=1048576
for rep in 1 2; do
run "H$H stock(q104 sf1)" OCD_KERNEL_FORM=wide
for v in ${VARIANTS:-q136sf1 q136sf3 q136sf5 q136sf0 q168sf3 q168sf5 q168sf0}; do
  run "H$H $v" OCD_B200_LIB=scratch/libocd_$v.so OCD_KERNEL_FORM=wide
done
done
