#!/bin/bash
# round-2 GPU call M: host-entry chunk size against the wave quantum of kernel 1 (592 CTAs = 757.8 profiles at L=100)
O=gpurun_out/r02m; mkdir -p $O
export MWRT_QUIET=1
for c in 0 757 750 1135 1514 1500 2271 3030; do
  for s in 2 3; do
    if [ "$c" == "0" ]; then echo -n "chunk default streams=$s: " >> $O/e2e_chunk.log; MWRT_HOST_STREAMS=$s python tools/e2e_time.py >> $O/e2e_chunk.log 2>&1
    else echo -n "chunk $c streams=$s: " >> $O/e2e_chunk.log; MWRT_HOST_CHUNK=$c MWRT_HOST_STREAMS=$s python tools/e2e_time.py >> $O/e2e_chunk.log 2>&1; fi
  done
done
cat $O/e2e_chunk.log
