#!/bin/bash
# round-2 GPU call L: band-pass channels per CTA (config 3), timings of the N1/N3/N4 kernels.
O=gpurun_out/r02l; mkdir -p $O
export MWRT_QUIET=1
for cpb in 7 4 3 2 1; do
  echo "== MWRT_BP_CPB=$cpb" >> $O/bp_cpb.log
  MWRT_BP_CPB=$cpb python tools/prof_target2.py c3 2000 5 >> $O/bp_cpb.log 2>&1
done
cat $O/bp_cpb.log
python tools/time_next_rows.py > $O/next_rows.log 2>&1; cat $O/next_rows.log
