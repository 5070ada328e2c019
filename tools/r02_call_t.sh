#!/bin/bash
# round-2 GPU call T: far lines on a coarse Chebyshev grid + per-block scratch built once (kernel 1) --
# timing against the full evaluation (MWRT_ABS_FAR=0) and the far-less build, checksums, GPU test suite
O=gpurun_out/r02t${1:-}; mkdir -p $O
export MWRT_QUIET=1
{
for v in main prenofar; do
  if [ "$v" == "main" ]; then unset MWRT_LIB; else export MWRT_LIB=$PWD/build_variants/libmwrt_$v.so; fi
  for far in 0 1; do
    if [ "$v" != "main" ] && [ "$far" == "1" ]; then continue; fi
    echo "== $v far=$far"
    MWRT_ABS_FAR=$far python tools/abs_time.py
    MWRT_ABS_FAR=$far ABS_TIME_BANDS=1 python tools/abs_time.py 2000 100 700
    MWRT_ABS_FAR=$far python tools/abs_time.py 300 150 4001
  done
done
unset MWRT_LIB
for w in c2 c2f32 c3; do python tools/prof_target2.py $w; done
} > $O/timings.log 2>&1
tail -n 40 $O/timings.log
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 5 $O/pytest_gpu.log
