#!/bin/bash
# round-2 GPU call U: where does the far-line build lose time against the round-2 base when nothing is far?
O=gpurun_out/r02u${1:-}; mkdir -p $O
export MWRT_QUIET=1
run() {  # lib far
  echo "== $1 far=$2"
  MWRT_LIB=$PWD/build_variants/libmwrt_$1.so MWRT_ABS_FAR=$2 python tools/abs_time.py
  MWRT_LIB=$PWD/build_variants/libmwrt_$1.so MWRT_ABS_FAR=$2 ABS_TIME_BANDS=1 python tools/abs_time.py 2000 100 700
}
{
run ident 0
run nofarHO 0
run noprol 0
run nofarH 0
run nofarO 0
run nofarH 1
run nofarO 1
} > $O/timings.log 2>&1
cat $O/timings.log
