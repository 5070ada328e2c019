#!/bin/bash
# compute-sanitizer on the smallest case that takes every kernel (tools/sanitize_target.py).
# NOTE: on this GPU pool the tool answers "closed on this pool and stays closed" (profiles/r02_sanitizer_closed.log).
# ONE tool per gpurun call (B200_PROFILING.md): tools/r02_call_sanitize.sh memcheck|racecheck|initcheck|synccheck
T=${1:-memcheck}
O=gpurun_out/r02s; mkdir -p $O
export MWRT_QUIET=1
python tools/sanitize_target.py > $O/plain.log 2>&1 || { echo "plain run failed"; tail -20 $O/plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool $T --print-limit 50 python tools/sanitize_target.py > $O/sanitizer_$T.log 2>&1
echo "compute-sanitizer $T rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|checksum|done|Error|error" $O/sanitizer_$T.log | tail -40
