#!/bin/bash
# round-2 GPU call X (final evidence on the final build): GPU tests, smoke, bench line, launch list of the
# bench command, ncu --set full of config 2's kernel 1
O=gpurun_out/r02x; mkdir -p $O
export MWRT_QUIET=1
timeout 300 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -2 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
python tools/prof_target2.py c2 10000 2 > $O/plain_c2.log 2>&1 &&
ncu --set full --clock-control none -k regex:'k_absorption' -s 1 -c 1 -o $O/prof_c2 python tools/prof_target2.py c2 10000 2 > $O/ncu_c2.log 2>&1
python tools/ncu_summary.py $O/prof_c2.ncu-rep > $O/prof_c2_abs.txt 2>&1; rm -f $O/*.ncu-rep
grep -E "time_duration|pipe_fp64_cycles|issue_active|dram__bytes|warps_active" $O/prof_c2_abs.txt
python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/ncu_bench.log 2>&1
{ MWRT_ABS_CARVEOUT=100 python tools/abs_time.py; MWRT_ABS_CARVEOUT=100 ABS_TIME_BANDS=1 python tools/abs_time.py 2000 100 700; python tools/abs_time.py; } > $O/carveout100.log 2>&1; cat $O/carveout100.log
python tools/launch_list.py $O/launches.csv > $O/launch_list.txt 2>&1; head -12 $O/launch_list.txt
python - <<PY
import json
d = json.loads(open("$O/bench_n1.json").read().strip().splitlines()[-1])
print("bench value %.4e ms/step %.4f e2e %.3e roofline frac %.3f k_ms %.4f dtb %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d.get("max_abs_dtb_vs_oracle_K")))
PY
