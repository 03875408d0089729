#!/bin/bash
cd /root/repo
cp l4dc-mpc-ocd_b200/libocd_b200.so /tmp/libocd_orig.so
for r in 128 112 168; do
  if [ $r = 128 ]; then cp /tmp/libocd_orig.so l4dc-mpc-ocd_b200/libocd_b200.so; else cp scratch/libocd_w$r.so l4dc-mpc-ocd_b200/libocd_b200.so; fi
  for pt in "15 2 262144" "15 6 262144" "50 2 65536" "50 6 65536" "5 6 1048576" "5 4 1048576" "5 3 1048576"; do
    set -- $pt
    echo -n "regs=$r wide "; OCD_KERNEL_FORM=wide python scripts/profile_point.py --H $1 --C $2 --B $3 --reps 3
  done
done
cp /tmp/libocd_orig.so l4dc-mpc-ocd_b200/libocd_b200.so
for pt in "5 6 1048576" "5 4 1048576" "5 3 1048576"; do set -- $pt; echo -n "throughput "; OCD_KERNEL_FORM=throughput python scripts/profile_point.py --H $1 --C $2 --B $3 --reps 3; done
