#!/bin/bash
cd /root/repo
cp l4dc-mpc-ocd_b200/libocd_b200.so /tmp/libocd_orig.so
for r in 128 168; do
  cp scratch/libocd_sf$r.so l4dc-mpc-ocd_b200/libocd_b200.so
  for pt in "5 6 1048576" "5 4 1048576" "5 6 65536"; do set -- $pt; echo -n "sf regs=$r wide "; OCD_KERNEL_FORM=wide python scripts/profile_point.py --H $1 --C $2 --B $3 --reps 3; done
done
cp /tmp/libocd_orig.so l4dc-mpc-ocd_b200/libocd_b200.so
for pt in "5 6 1048576" "5 4 1048576" "5 6 65536"; do set -- $pt; echo -n "throughput "; OCD_KERNEL_FORM=throughput python scripts/profile_point.py --H $1 --C $2 --B $3 --reps 3; done
