#!/usr/bin/env python3
"""Development helper: time kernel 1 alone (config 2) -- `python tools/abs_time.py [P] [L] [F-list]`."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import mwr_fast_forward_operators_and_lbls_b200 as pkg
from mwr_fast_forward_operators_and_lbls_b200 import _lib
from mwr_fast_forward_operators_and_lbls_b200.forward import _ptr, _stream
P = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 100
nf = int(sys.argv[3]) if len(sys.argv) > 3 else 14
frq = pkg.HATPRO_FRQ if nf == 14 else np.linspace(20.0, 60.0, nf)
if os.environ.get("ABS_TIME_BANDS") and nf % 14 == 0:   # config 3's sub-band frequencies
    frq = pkg.band_frequencies(nf // 14)
d = pkg.synth_profiles(P, L, seed=1234)
dev = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
e = pkg.vapor(dev["t"], dev["rh"])[0]
lib = _lib.load()
f = torch.from_numpy(np.ascontiguousarray(frq)).cuda()
awet = torch.empty((P, nf, L), dtype=torch.float64, device="cuda"); adry = torch.empty_like(awet)
m = pkg.get_model(sys.argv[4] if len(sys.argv) > 4 else "R24")
plan = pkg.Plan(sys.argv[4] if len(sys.argv) > 4 else "R24", L, frq, pkg.HATPRO_ELEV[:1])
def run():   # the plan's stage call: per-block tables cached, exactly the forward pass's launch
    if os.environ.get("ABS_TIME_STAGE"):
        lib.mwrt_absorption_f64(m.handle, P, L, nf, _ptr(dev["p"]), _ptr(dev["t"]), _ptr(e), None, None, _ptr(f), _ptr(awet), _ptr(adry), None, None, None, _stream())
    else:
        plan.absorption(dev["p"], dev["t"], e, out=(awet, adry))
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(20):
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print("absorption P=%d L=%d F=%d far=%s: %.4f ms (min %.4f) checksum %.10e" % (P, L, nf, os.environ.get("MWRT_ABS_FAR", "1") + ("/bands" if os.environ.get("ABS_TIME_BANDS") else ""), float(np.median(ts)), min(ts), float(awet.sum() + adry.sum())))
