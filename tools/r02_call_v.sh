#!/bin/bash
# round-2 GPU call V: far lines on the coarse grid in the wide-block kernels + per-block scratch built once
# per plan -- tests, kernel-1 timings (plan stage call), configs 2/3/4, bench line, ncu of config 3's kernel 1
O=gpurun_out/r02v${1:-}; mkdir -p $O
export MWRT_QUIET=1
timeout 600 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
{
python tools/abs_time.py
for far in 1 0; do
  MWRT_ABS_FAR=$far ABS_TIME_BANDS=1 python tools/abs_time.py 2000 100 700
  MWRT_ABS_FAR=$far python tools/abs_time.py 300 150 4001
done
for t in c2 c2f32 c2rt; do python tools/prof_target2.py $t 10000 8; done
python tools/prof_target2.py c3 2000 5
python tools/prof_target2.py c4 500 3
} > $O/timings.log 2>&1
cat $O/timings.log
python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
ncu --set full --clock-control none -k regex:'k_absorption' -s 1 -c 1 -o $O/prof_c3 python tools/prof_target2.py c3 2000 2 > $O/ncu_c3.log 2>&1
python tools/ncu_summary.py $O/prof_c3.ncu-rep > $O/prof_c3_abs.txt 2>&1; rm -f $O/*.ncu-rep
grep -E "time_duration|pipe_fp64_cycles|issue_active|dram__bytes|registers|warps_active" $O/prof_c3_abs.txt
python - <<PY
import json
d = json.loads(open("$O/bench_n1.json").read().strip().splitlines()[-1])
print("bench value %.4e ms/step %.4f e2e %.3e roofline frac %.3f k_ms %.4f dtb %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d.get("max_abs_dtb_vs_oracle_K")))
for k, v in (d.get("extra", {}).get("configs", {}) or {}).items():
    print(k, {kk: v[kk] for kk in ("ms_per_pass", "ms_kernels_only") if kk in v})
PY
