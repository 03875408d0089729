#!/bin/bash
# do out-of-phase warps help kernels of a few waves?  (GPU box)
cd "$(dirname "$0")/../.."
run() { echo -n "$1 "; shift; env "$@" python scripts/profile_point.py --H $H --C 2 --B $B --reps 20; }
H=15
for B in 65536 131072 1048576; do
  run "H15 B=$B stock" OCD_KERNEL_FORM=wide
  for v in st300 st1000 st5k; do run "H15 B=$B $v" OCD_B200_LIB=scratch/libocd_$v.so OCD_KERNEL_FORM=wide; done
done
H=5
for B in 65536 131072 1048576; do
  run "H5 B=$B stock" OCD_KERNEL_FORM=wide
  for v in h5st100 h5st400; do run "H5 B=$B $v" OCD_B200_LIB=scratch/libocd_$v.so OCD_KERNEL_FORM=wide; done
done
