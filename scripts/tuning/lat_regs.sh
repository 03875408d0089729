#!/bin/bash
# experiment: the latency form of k_solve<5,1,3> under different register caps, at large batch sizes
cd /root/repo
cp l4dc-mpc-ocd_b200/libocd_b200.so /tmp/libocd_orig.so
for r in orig 96 112 128; do
  if [ $r = orig ]; then cp /tmp/libocd_orig.so l4dc-mpc-ocd_b200/libocd_b200.so; else cp scratch/libocd_lat$r.so l4dc-mpc-ocd_b200/libocd_b200.so; fi
  for form in throughput latency; do
    for B in 262144 1048576; do
      echo -n "regs=$r form=$form "; OCD_KERNEL_FORM=$form python scripts/profile_point.py --H 5 --C 2 --B $B --reps 5
    done
  done
done
cp /tmp/libocd_orig.so l4dc-mpc-ocd_b200/libocd_b200.so
