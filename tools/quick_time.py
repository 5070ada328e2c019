#!/usr/bin/env python3
"""Development helper: time the stage kernels of config 2 with CUDA events (not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import mwr_fast_forward_operators_and_lbls_b200 as pkg

P = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 100
mdl = sys.argv[3] if len(sys.argv) > 3 else "R24"
d = pkg.synth_profiles(P, L, seed=1234)
dev = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
fl, mhz = pkg.measure_dfma_peak()
print("DFMA peak %.2f TFLOP/s, clock est %.0f MHz" % (fl / 1e12, mhz))
plan = pkg.Plan(mdl, L, pkg.HATPRO_FRQ, pkg.HATPRO_ELEV)
for it in range(3):
    out = plan.forward(dev["z"], dev["p"], dev["t"], dev["rh"])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for it in range(10):
    e0.record(); out = plan.forward(dev["z"], dev["p"], dev["t"], dev["rh"], out=out); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
ntb = P * 14 * 10
print("forward: %.3f ms  -> %.3e TB/s ; min %.3f" % (ms, ntb / ms * 1e3, min(ts)))
# stages
e = pkg.vapor(dev["t"], dev["rh"])[0]
from mwr_fast_forward_operators_and_lbls_b200 import _lib
from mwr_fast_forward_operators_and_lbls_b200.forward import _ptr, _stream
lib = _lib.load()
f = torch.from_numpy(pkg.HATPRO_FRQ).cuda(); a = torch.from_numpy(pkg.HATPRO_ELEV).cuda()
awet = torch.empty((P, 14, L), dtype=torch.float64, device="cuda"); adry = torch.empty_like(awet)
ds = torch.empty((P, 10, L), dtype=torch.float64, device="cuda")
outs = [torch.empty((P, 14, 10), dtype=torch.float64, device="cuda") for _ in range(7)]
m = pkg.get_model(mdl)
def t_abs():
    lib.mwrt_absorption_f64(m.handle, P, L, 14, _ptr(dev["p"]), _ptr(dev["t"]), _ptr(e), None, None, _ptr(f), _ptr(awet), _ptr(adry), None, None, None, _stream())
def t_ds():
    lib.mwrt_slant_path_f64(P, L, 10, _ptr(dev["z"]), _ptr(dev["p"]), _ptr(dev["t"]), _ptr(e), _ptr(a), 0, _ptr(ds), None, _stream())
def t_rte():
    lib.mwrt_rte_f64(P, L, 14, 10, _ptr(dev["t"]), _ptr(awet), _ptr(adry), None, None, _ptr(ds), _ptr(f), 0, 1.0, *[_ptr(o) for o in outs], None, 0, None, _stream())
for name, fn in (("absorption", t_abs), ("slant", t_ds), ("rte", t_rte)):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(10):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print("%-10s %.3f ms (min %.3f)" % (name, float(np.median(ts)), min(ts)))
# host path
hp = pkg.pinned_fields(list(d.keys()), d["z"].shape)
for k in d: hp[k][...] = d[k]
ho = pkg.pinned_fields(pkg.FIELDS, (P, 14, 10))
ho["status"] = torch.empty((P,), dtype=torch.int32).pin_memory().numpy()
for _ in range(3): plan.forward_host(hp["z"], hp["p"], hp["t"], hp["rh"], out=ho)
ts = []
for _ in range(10):
    t0 = time.perf_counter(); plan.forward_host(hp["z"], hp["p"], hp["t"], hp["rh"], out=ho); ts.append((time.perf_counter() - t0) * 1e3)
print("host e2e: %.3f ms (min %.3f) -> %.3e TB/s" % (float(np.median(ts)), min(ts), ntb / float(np.median(ts)) * 1e3))
