#!/bin/bash
# round-2 GPU call A: (1) offline probe for real pyrtlib on the GPU box (no network fetch: installing from a
# public index is not permitted in this environment), (2) timings of the paths round 1 never measured
# (ray tracing, cloudy kernels), (3) ncu --set full of those kernels.
O=gpurun_out/r02a; mkdir -p $O
{
  echo "== import pyrtlib"; python -c "import pyrtlib; print(pyrtlib.__version__)"
  echo "== local wheelhouse"; ls /opt/wheelhouse 2>/dev/null | grep -i -E 'pyrtlib|xarray|netcdf' || echo "none"
  echo "== find"; find / -iname '*pyrtlib*' -not -path '/proc/*' 2>/dev/null | head
  echo "== probe"; python baseline/probe_pyrtlib.py $O/pyrtlib_goldens.json
} > $O/probe_pyrtlib.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.csv
python tools/prof_target2.py c2 10000 8 > $O/time_c2.log 2>&1
python tools/prof_target2.py c2rt 10000 8 > $O/time_c2rt.log 2>&1
python tools/prof_target2.py c3 2000 5 > $O/time_c3.log 2>&1
python tools/run_configs.py 3 > $O/configs3.jsonl 2> $O/configs3.err
python tools/prof_target2.py c3 1000 2 > $O/plain_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_absorption|k_rte|k_bandpass' -s 24 -c 3 -o $O/prof_c3 python tools/prof_target2.py c3 1000 2 > $O/ncu_c3.log 2>&1
python tools/prof_target2.py c2rt 10000 2 > $O/plain_c2rt.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_slant|k_rte' -s 2 -c 2 -o $O/prof_c2rt python tools/prof_target2.py c2rt 10000 2 > $O/ncu_c2rt.log 2>&1
ls -la $O
