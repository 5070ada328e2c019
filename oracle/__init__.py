"""CPU oracle for the PyRTlib `TbCloudRTE` path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

What this is
------------
A plain fp64 restatement (pure Python / NumPy) of the algorithm the reference drives through
`pyrtlib.tb_spectrum.TbCloudRTE` (`/root/reference/python_src/proc/PyRTlib_processing.py:123-127`,
`python_src/proc/old_processing.py:556-565`).  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import it -- and only as the checker or the
timed CPU arm.  Nothing under `mwr_fast_forward_operators_and_lbls_b200/` imports it.

PARITY UNPINNED
---------------
The arithmetic of this path lives in the third-party package `pyrtlib==1.1.1`
(`/root/reference/requirements.txt:278`; put on `sys.path` at `PyRTlib_processing.py:26`).  Its
source is not under `/root/reference`, it is not installable offline (no wheel in /opt/wheelhouse),
and the reference has no tests, golden vectors or stored outputs for the path.
This restatement follows the published algorithm (Rosenkranz's public-domain `abh2o`/`o2abs`/`absn2`/
`abliq12` routines and the NOAA/ETL `RTE` routines as re-expressed by PyRTlib) from memory; it has been
checked against physics invariants and against a second, independently written vectorised restatement
(`oracle/pyrtlib_vec.py`), NOT against PyRTlib output.  `baseline/probe_pyrtlib.py` records true
goldens whenever a box has `pyrtlib` installed.

Modules
-------
model_tables    loads the shared model descriptors (data, not code)
pyrtlib_scalar  PyRTlib-shaped scalar loops (one object per call, per-angle recomputation)
pyrtlib_vec     vectorised NumPy form used for parity at sizes of 10^2..10^3 profiles
"""
