#!/bin/bash
# round-2 GPU call J: software-pipelined pair loops of kernel 1 (variants built by tools/build_variant.sh)
O=gpurun_out/r02j; mkdir -p $O
export MWRT_QUIET=1
for v in ${VARIANTS:-base o2p4 o2p3 hop4 hop3}; do
  echo "== $v" >> $O/variants.log
  MWRT_LIB=$PWD/build_variants/libmwrt_$v.so python tools/abs_time.py >> $O/variants.log 2>&1
  MWRT_LIB=$PWD/build_variants/libmwrt_$v.so python tools/abs_time.py 2000 100 700 >> $O/variants.log 2>&1
  MWRT_LIB=$PWD/build_variants/libmwrt_$v.so python tools/abs_time.py 300 150 4001 >> $O/variants.log 2>&1
  MWRT_LIB=$PWD/build_variants/libmwrt_$v.so python tools/prof_target2.py c2 10000 8 >> $O/variants.log 2>&1
done
cat $O/variants.log
