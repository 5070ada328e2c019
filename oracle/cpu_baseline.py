"""CPU arm of the benchmark: the PyRTlib-shaped oracle timed on the host cores.  TEST/BENCH
INFRASTRUCTURE (see oracle/__init__.py) -- `bench.py` runs it as a subprocess for `cpu_baseline` and
for `--impl reference`; nothing in the product imports it.

It reproduces the reference's call pattern exactly (`PyRTlib_processing.py:99-127`): one
`TbCloudRTE` object per (profile, elevation) with ONE angle, absorption recomputed for every angle,
`multiprocessing.Pool` over the host cores (the reference itself is single-threaded; all cores is the
generous reading BASELINE.md section 3 asks for).  kind = "port": real `pyrtlib` is not installable
offline; if `import pyrtlib` ever succeeds on the box it is timed instead and kind = "reference".

    python -m oracle.cpu_baseline --tasks 64 --levels 100 --model R24 [--steps K --warmup W]
prints one JSON line: {"value": TB/s, "unit": ..., "cores": C, "kind": ..., "sample": ..., "step_s": [...]}
"""
import argparse
import importlib.util
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _load_synth():
    """synth.py is NumPy-only; load it by path so the worker never imports torch / the CUDA package."""
    path = os.path.join(ROOT, "mwr_fast_forward_operators_and_lbls_b200", "synth.py")
    spec = importlib.util.spec_from_file_location("_mwrt_synth", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_G = {}


def _init(profiles, frq, model, use_pyrtlib):
    _G["d"] = profiles
    _G["frq"] = frq
    _G["model"] = model
    _G["pyrtlib"] = use_pyrtlib


def _task(arg):
    """One `execute()` exactly as the reference issues it: one profile, ONE elevation."""
    i, angle = arg
    d = _G["d"]
    if _G["pyrtlib"]:
        from pyrtlib.tb_spectrum import TbCloudRTE  # pragma: no cover (not installable offline)
    else:
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        from oracle.pyrtlib_scalar import TbCloudRTE
    rte = TbCloudRTE(d["z"][i], d["p"][i], d["t"][i], d["rh"][i], _G["frq"], np.array([angle]))
    rte.init_absmdl(_G["model"])
    rte.satellite = False
    df = rte.execute()
    return np.asarray(df["tbtotal"])


def have_pyrtlib():
    try:
        import pyrtlib  # noqa: F401
        return True
    except Exception:
        return False


def run(tasks, levels, model, steps=1, warmup=0, cores=None, seed=1234):
    synth = _load_synth()
    cores = cores or os.cpu_count() or 1
    angles = synth.HATPRO_ELEV
    nprof = max(1, -(-tasks // len(angles)))
    d = synth.synth_profiles(nprof, levels, seed=seed)
    work = [(i, float(a)) for i in range(nprof) for a in angles][:tasks]
    use_real = have_pyrtlib()
    step_s = []
    with mp.Pool(cores, initializer=_init, initargs=(d, synth.HATPRO_FRQ, model, use_real)) as pool:
        for s in range(warmup + steps):
            t0 = time.perf_counter()
            res = pool.map(_task, work, chunksize=1)
            dt = time.perf_counter() - t0
            if s >= warmup:
                step_s.append(dt)
    ntb = len(work) * len(synth.HATPRO_FRQ)
    total = sum(step_s)
    return dict(value=ntb * len(step_s) / total, unit="profile*channel*angle TB/s", cores=cores,
                kind="reference" if use_real else "port",
                sample="%d execute() calls per step (%d profiles x up to %d elevations, one angle per "
                       "TbCloudRTE object as PyRTlib_processing.py:106), %d levels, 14 channels, model "
                       "%s, %d step(s)" % (len(work), nprof, len(angles), levels, model, len(step_s)),
                step_s=step_s, s_per_execute=total / len(step_s) / len(work) * min(cores, len(work)),
                tb_per_step=ntb, checksum=float(np.sum(res)))


def parity_sample(path, nprof, levels, model, seed=4321):
    """The checker half of bench.py's cpu_baseline leg: the vectorised oracle on `nprof` synthetic
    profiles (seeded, so bench.py feeds the SAME profiles to the GPU), all channels and elevations;
    written to `path` (.npz) for bench.py to take max |dTB| against, outside any timed region."""
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import pyrtlib_vec
    synth = _load_synth()
    d = synth.synth_profiles(nprof, levels, seed=seed)
    ref = pyrtlib_vec.simulate(d["z"], d["p"], d["t"], d["rh"], synth.HATPRO_FRQ, synth.HATPRO_ELEV, model)
    np.savez(path, seed=seed, nprof=nprof, levels=levels, model=model,
             **{k: ref[k] for k in ("tbtotal", "tbatm", "tmr", "tauwet", "taudry")})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tasks", type=int, default=0, help="execute() calls per step (0: calibrate)")
    ap.add_argument("--levels", type=int, default=100)
    ap.add_argument("--model", default="R24")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--cores", type=int, default=0)
    ap.add_argument("--budget-s", type=float, default=15.0, help="target CPU wall time for all steps")
    ap.add_argument("--parity-out", default="", help="also write the oracle's outputs for a seeded sample (.npz)")
    ap.add_argument("--parity-profiles", type=int, default=256)
    a = ap.parse_args()
    if a.parity_out:
        parity_sample(a.parity_out, a.parity_profiles, a.levels, a.model)
    cores = a.cores or os.cpu_count() or 1
    tasks = a.tasks
    if tasks <= 0:
        # calibrate: one execute() on one core, then size the sample for the budget
        t0 = time.perf_counter()
        _init(_load_synth().synth_profiles(1, a.levels, seed=1), _load_synth().HATPRO_FRQ, a.model,
              have_pyrtlib())
        _task((0, 90.0))
        t_exec = time.perf_counter() - t0
        per_step = a.budget_s / max(1, a.steps + a.warmup)
        tasks = int(max(cores, min(20000, per_step / t_exec * cores)))
        tasks = max(cores, (tasks // cores) * cores)
    print(json.dumps(run(tasks, a.levels, a.model, a.steps, a.warmup, cores)))


if __name__ == "__main__":
    main()
