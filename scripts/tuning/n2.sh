#!/bin/bash
cd /root/repo
cp l4dc-mpc-ocd_b200/libocd_b200.so /tmp/libocd_orig.so
for lib in orig n2; do
  if [ $lib = orig ]; then cp /tmp/libocd_orig.so l4dc-mpc-ocd_b200/libocd_b200.so; else cp scratch/libocd_n2_128.so l4dc-mpc-ocd_b200/libocd_b200.so; fi
  for pt in "5 3 1048576" "5 3 262144" "5 3 65536"; do set -- $pt; echo -n "$lib "; python scripts/profile_point.py --H $1 --C $2 --B $3 --reps 3; done
done
cp /tmp/libocd_orig.so l4dc-mpc-ocd_b200/libocd_b200.so
