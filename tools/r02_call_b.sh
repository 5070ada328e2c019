#!/bin/bash
# round-2 GPU call B: GPU tests of the reworked kernels, timings, first bench line, ncu of the new kernels.
O=gpurun_out/r02b; mkdir -p $O
export MWRT_QUIET=1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
for t in c2 c2rt c2f32; do python tools/prof_target2.py $t 10000 8 > $O/time_$t.log 2>&1; done
python tools/prof_target2.py c3 2000 5 > $O/time_c3.log 2>&1
MWRT_ABS_WIDE=1 python tools/run_configs.py 3 > $O/configs3_wide1.jsonl 2> $O/configs3_wide1.err
MWRT_ABS_WIDE=2 python tools/run_configs.py 3 > $O/configs3_wide2.jsonl 2> $O/configs3_wide2.err
MWRT_ABS_WIDE=1 python tools/run_configs.py 4 --scale 0.2 > $O/configs4_wide1.jsonl 2> $O/configs4_wide1.err
MWRT_ABS_WIDE=2 python tools/run_configs.py 4 --scale 0.2 > $O/configs4_wide2.jsonl 2> $O/configs4_wide2.err
python tools/run_configs.py 5 > $O/configs5.jsonl 2> $O/configs5.err
cat $O/time_*.log $O/configs*.jsonl
python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
python tools/prof_target2.py c3 1000 2 > $O/plain_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_absorption|k_rte' -s 4 -c 2 -o $O/prof_c3 python tools/prof_target2.py c3 1000 2 > $O/ncu_c3.log 2>&1
python tools/prof_target2.py c2rt 10000 2 > $O/plain_c2rt.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_slant|k_rte' -s 2 -c 2 -o $O/prof_c2rt python tools/prof_target2.py c2rt 10000 2 > $O/ncu_c2rt.log 2>&1
python tools/prof_target2.py c2f32 10000 2 > $O/plain_c2f32.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_absorption' -s 1 -c 1 -o $O/prof_c2f32 python tools/prof_target2.py c2f32 10000 2 > $O/ncu_c2f32.log 2>&1
python tools/prof_target2.py c2 10000 2 > $O/plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_absorption|k_rte' -s 2 -c 2 -o $O/prof_c2 python tools/prof_target2.py c2 10000 2 > $O/ncu_c2.log 2>&1
ls -la $O
