#!/bin/bash
# Q kernel (H = 15): larger blocks and a barrier per iteration against instruction-fetch stalls (GPU box)
cd "$(dirname "$0")/../.."
run() { echo -n "$1 "; shift; env "$@" python scripts/profile_point.py --H 15 --C 2 --B 1048576 --reps 4; }
run "H15 stock(P=32, no barrier)" OCD_KERNEL_FORM=wide
for v in qsync32 qsync64 qp160 qsync160; do run "H15 $v" OCD_B200_LIB=scratch/libocd_$v.so OCD_KERNEL_FORM=wide; done
