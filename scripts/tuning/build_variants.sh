#!/bin/bash
# Build variant libraries for A/B measurements on the GPU box: the stock objects of csrc/build with ONE
# translation unit recompiled under extra -D flags.  Run here (nvcc cross-compiles), the .so files travel with gpurun.
#   scripts/tuning/build_variants.sh <name> <unit: solve_15_1_3 | inst_5_1_3_0 | ...> "<-D flags>"
# -> scratch/libocd_<name>.so   (load it with OCD_B200_LIB=scratch/libocd_<name>.so)
set -eu
NAME=$1; UNIT=$2; FLAGS=$3
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
CS=$ROOT/l4dc-mpc-ocd_b200/csrc
OUT=$ROOT/scratch/var/$NAME
mkdir -p "$OUT"
# objects without -lineinfo (a quarter of the size: the variant libraries must fit gpurun's snapshot limit)
NL=$ROOT/scratch/var/build_nl
[ -d "$NL" ] && [ "$NL" -nt "$CS/ocd_device.cuh" ] && [ "$NL" -nt "$CS/ocd_kernels.cuh" ] && [ "$NL" -nt "$CS/ocd_api.cu" ] || {
  rm -rf "$NL"; make -s -j8 -C "$CS" BUILD="$NL" OUT="$ROOT/scratch/var/libocd_nl.so" LINEINFO= >/dev/null; touch "$NL"; }
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -I$ROOT/include -I$CS -Xptxas -v"
case $UNIT in
  solve_*) IFS=_ read -r _ HT NO LT <<< "$UNIT"
           $NV $FLAGS -DOCD_SOLVE_ONLY -DOCD_PRECISE=0 -DOCD_HT=$HT -DOCD_NO=$NO -DOCD_LT=$LT -c $CS/ocd_inst.cu -o $OUT/$UNIT.o 2> $OUT/ptxas.log ;;
  inst_*)  IFS=_ read -r _ HT NO LT PR <<< "$UNIT"
           $NV $FLAGS -DOCD_PRECISE=$PR -DOCD_HT=$HT -DOCD_NO=$NO -DOCD_LT=$LT -c $CS/ocd_inst.cu -o $OUT/$UNIT.o 2> $OUT/ptxas.log ;;
  api)     $NV $FLAGS -c $CS/ocd_api.cu -o $OUT/api.o 2> $OUT/ptxas.log ;;
esac
OBJS=$(ls $NL/*.o | grep -v "/$UNIT.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/scratch/libocd_$NAME.so $OBJS $OUT/$UNIT.o
grep -E "spill|registers" $OUT/ptxas.log | paste - - | sed 's/ptxas info    ://g' | head -6
