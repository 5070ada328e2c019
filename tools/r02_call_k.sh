#!/bin/bash
# round-2 GPU call K: after adopting the O2 software pipeline for narrow blocks -- tests, kernel-1 timings, bench.
O=gpurun_out/r02k; mkdir -p $O
export MWRT_QUIET=1
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
python tools/abs_time.py > $O/timings.log 2>&1
python tools/abs_time.py 2000 100 700 >> $O/timings.log 2>&1
python tools/abs_time.py 300 150 4001 >> $O/timings.log 2>&1
for t in c2 c2rt c2f32; do python tools/prof_target2.py $t 10000 8; done >> $O/timings.log 2>&1
python tools/prof_target2.py c3 2000 5 >> $O/timings.log 2>&1; cat $O/timings.log
python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
python tools/prof_target2.py c2 10000 2 > $O/plain_c2.log 2>&1 &&
ncu --set full --clock-control none -k regex:'k_absorption' -s 1 -c 1 -o $O/prof_c2 python tools/prof_target2.py c2 10000 2 > $O/ncu_c2.log 2>&1
python tools/ncu_summary.py $O/prof_c2.ncu-rep > $O/prof_c2.txt 2>&1; rm -f $O/*.ncu-rep
grep -E "time_duration|pipe_fp64_cycles|issue_active|dram__bytes" $O/prof_c2.txt
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02k/bench_n1.json"))
print("bench value %.4e ms/step %.4f e2e %.3e roofline frac %.3f k_ms %.4f dtb %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["max_abs_dtb_vs_oracle_K"]))
PY
