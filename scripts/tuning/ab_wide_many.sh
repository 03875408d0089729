#!/bin/bash
# register cap of the wide form with 2 / 4 / 5 other cars at H = 5 (GPU box)
cd "$(dirname "$0")/../.."
run() { echo -n "$1 "; shift; env "$@" python scripts/profile_point.py --H 5 --C $C --B 1048576 --reps 5; }
for C in 3 5 6; do
  u=5_$((C-1))_3
  run "C=$C stock(168)" OCD_KERNEL_FORM=wide
  for r in 128 144 152; do run "C=$C regs=$r" OCD_B200_LIB=scratch/libocd_wm${r}_$u.so OCD_KERNEL_FORM=wide; done
done
