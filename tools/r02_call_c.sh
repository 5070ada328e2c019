#!/bin/bash
# round-2 GPU call C: full GPU test run, kernel-1 / kernel-3 tuning variants, bench line, ncu summaries
# (reports are condensed on the box: gpurun_out may carry 64 MiB back).
O=gpurun_out/r02c; mkdir -p $O
export MWRT_QUIET=1
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -15 $O/pytest_gpu.log
echo "== main" > $O/variants.log
python tools/abs_time.py >> $O/variants.log 2>&1
python tools/abs_time.py 2000 100 700 >> $O/variants.log 2>&1
python tools/prof_target2.py c2 10000 8 >> $O/variants.log 2>&1
python tools/prof_target2.py c2rt 10000 8 >> $O/variants.log 2>&1
for v in abs_g0 abs_g4 abs_g14 abs_m5 abs_m3; do
  echo "== $v" >> $O/variants.log
  MWRT_LIB=$PWD/build_variants/libmwrt_$v.so python tools/abs_time.py >> $O/variants.log 2>&1
  MWRT_LIB=$PWD/build_variants/libmwrt_$v.so python tools/abs_time.py 2000 100 700 >> $O/variants.log 2>&1
  MWRT_LIB=$PWD/build_variants/libmwrt_$v.so python tools/prof_target2.py c2 10000 8 >> $O/variants.log 2>&1
done
echo "== rte_g2" >> $O/variants.log
MWRT_LIB=$PWD/build_variants/libmwrt_rte_g2.so python tools/prof_target2.py c2rt 10000 8 >> $O/variants.log 2>&1
cat $O/variants.log
python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
python tools/prof_target2.py c3 1000 2 > $O/plain_c3.log 2>&1 &&
ncu --set full --clock-control none -k regex:'k_absorption|k_rte' -s 4 -c 2 -o $O/prof_c3 python tools/prof_target2.py c3 1000 2 > $O/ncu_c3.log 2>&1
python tools/prof_target2.py c2rt 10000 2 > $O/plain_c2rt.log 2>&1 &&
ncu --set full --clock-control none -k regex:'k_slant|k_rte' -s 2 -c 2 -o $O/prof_c2rt python tools/prof_target2.py c2rt 10000 2 > $O/ncu_c2rt.log 2>&1
python tools/prof_target2.py c2f32 10000 2 > $O/plain_c2f32.log 2>&1 &&
ncu --set full --clock-control none -k regex:'k_absorption' -s 1 -c 1 -o $O/prof_c2f32 python tools/prof_target2.py c2f32 10000 2 > $O/ncu_c2f32.log 2>&1
python tools/prof_target2.py c2 10000 2 > $O/plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_absorption' -s 1 -c 1 -o $O/prof_c2 python tools/prof_target2.py c2 10000 2 > $O/ncu_c2.log 2>&1
for r in c3 c2rt c2f32 c2; do python tools/ncu_summary.py $O/prof_$r.ncu-rep > $O/prof_$r.txt 2>&1; done
ncu -i $O/prof_c2.ncu-rep --page source --csv > $O/prof_c2_source.csv 2>/dev/null
rm -f $O/prof_c3.ncu-rep $O/prof_c2rt.ncu-rep $O/prof_c2f32.ncu-rep
gzip -f $O/prof_c2_source.csv
ls -la $O; du -sh $O
