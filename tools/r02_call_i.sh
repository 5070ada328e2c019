#!/bin/bash
# round-2 GPU call I (1 GPU): final state -- all GPU tests, the bench line, the ncu launch list of the bench
# command and `ncu --set full` summaries of the hot kernels (condensed on the box).
O=gpurun_out/r02i; mkdir -p $O
export MWRT_QUIET=1
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench ref rc=$?"
for t in c2 c2rt c2f32; do python tools/prof_target2.py $t 10000 8; done > $O/timings.log 2>&1
python tools/prof_target2.py c3 2000 5 >> $O/timings.log 2>&1; cat $O/timings.log
python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > $O/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/bench_launches.csv python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_bench.log 2>&1
python tools/launch_list.py $O/bench_launches.csv > $O/bench_launch_list.txt 2>&1; cat $O/bench_launch_list.txt
python tools/prof_target2.py c2 10000 2 > $O/plain_c2.log 2>&1 &&
ncu --set full --clock-control none -k regex:'k_absorption|k_rte' -s 2 -c 2 -o $O/prof_c2 python tools/prof_target2.py c2 10000 2 > $O/ncu_c2.log 2>&1
python tools/prof_target2.py c2rt 10000 2 > $O/plain_c2rt.log 2>&1 &&
ncu --set full --clock-control none -k regex:'k_slant|k_rte' -s 2 -c 2 -o $O/prof_c2rt python tools/prof_target2.py c2rt 10000 2 > $O/ncu_c2rt.log 2>&1
python tools/prof_target2.py c3 1000 2 > $O/plain_c3.log 2>&1 &&
ncu --set full --clock-control none -k regex:'k_absorption|k_rte' -s 4 -c 2 -o $O/prof_c3 python tools/prof_target2.py c3 1000 2 > $O/ncu_c3.log 2>&1
python tools/prof_target2.py c2f32 10000 2 > $O/plain_c2f32.log 2>&1 &&
ncu --set full --clock-control none -k regex:'k_absorption' -s 1 -c 1 -o $O/prof_c2f32 python tools/prof_target2.py c2f32 10000 2 > $O/ncu_c2f32.log 2>&1
for r in c2 c2rt c3 c2f32; do python tools/ncu_summary.py $O/prof_$r.ncu-rep > $O/prof_$r.txt 2>&1; done
rm -f $O/*.ncu-rep
ls -la $O; du -sh $O
