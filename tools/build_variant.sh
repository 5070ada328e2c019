#!/bin/bash
# build_variants/libmwrt_<name>.so = the in-tree objects with ONE translation unit recompiled with extra
# macros (tuning experiments; selected at run time with MWRT_LIB=<path>).
#   tools/build_variant.sh <name> <tu: absorption|rte> [-DMACRO=..]...
set -e
name=$1; tu=$2; shift 2
R=$(cd "$(dirname "$0")/.." && pwd); C=$R/mwr_fast_forward_operators_and_lbls_b200/csrc; V=$R/build_variants
mkdir -p $V
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v --fmad=true "$@" \
  -c $C/mwrt_$tu.cu -o $V/${tu}_$name.o > $V/${tu}_$name.log 2>&1
objs=""
for t in api absorption rte rte_fanout post peak push; do
  if [ "$t" == "$tu" ]; then objs="$objs $V/${tu}_$name.o"; else objs="$objs $C/mwrt_$t.o"; fi
done
nvcc -shared -o $V/libmwrt_$name.so $objs -gencode arch=compute_100a,code=sm_100a -cudart static
echo built $V/libmwrt_$name.so
