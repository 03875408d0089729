#!/bin/bash
cd /root/repo
cp l4dc-mpc-ocd_b200/libocd_b200.so /tmp/libocd_orig.so
for lib in orig seg4 seg6; do
  if [ $lib = orig ]; then cp /tmp/libocd_orig.so l4dc-mpc-ocd_b200/libocd_b200.so; else cp scratch/libocd_$lib.so l4dc-mpc-ocd_b200/libocd_b200.so; fi
  for pt in "15 2 262144" "50 2 65536" "15 6 262144" "50 6 65536" "16 2 262144"; do set -- $pt; echo -n "$lib "; python scripts/profile_point.py --H $1 --C $2 --B $3 --reps 3; done
done
cp /tmp/libocd_orig.so l4dc-mpc-ocd_b200/libocd_b200.so
