#!/bin/bash
# A/B of the world-tile staging: per-thread LDG (stock) vs 1-D TMA bulk copies + mbarrier (-DOCD_TMA_STAGE variant of the
# bench kernel's translation unit, scripts/tuning/build_variants.sh tma inst_5_1_3_0 "-DOCD_TMA_STAGE").  GPU box.
cd "$(dirname "$0")/../.."
for rep in 1 2; do
for pt in "48 400" "4096 200" "1048576 10"; do set -- $pt
  echo -n "B=$1 stock(tma) "; python scripts/profile_point.py --H 5 --C 2 --B $1 --reps $2
  echo -n "B=$1 ldgv "; OCD_B200_LIB=scratch/libocd_ldg.so python scripts/profile_point.py --H 5 --C 2 --B $1 --reps $2
done; done
