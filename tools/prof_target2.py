#!/usr/bin/env python3
"""ncu / timing target for the configurations that are not the bench line:

    python tools/prof_target2.py c2    [P] [iters]   config 2, plane-parallel (clear sky)
    python tools/prof_target2.py c2rt  [P] [iters]   config 2 with ray_tracing=True (k_slant_path + general k_rte)
    python tools/prof_target2.py c3    [P] [iters]   config 3 share: liquid cloud, 700 sub-frequencies -> 14 channels
    python tools/prof_target2.py c4    [P] [iters]   config 4 share: 4001 frequencies x 150 levels, per-layer tau
    python tools/prof_target2.py c2f32 [P] [iters]   config 2 with the fp32 line-shape mode

Prints the CUDA-event time per forward (median) so the same command serves as the plain run that must
exit 0 before ncu wraps it.
"""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import mwr_fast_forward_operators_and_lbls_b200 as pkg

what = sys.argv[1] if len(sys.argv) > 1 else "c2"
defaults = {"c2": 10000, "c2rt": 10000, "c3": 2000, "c4": 500, "c2f32": 10000}
P = int(sys.argv[2]) if len(sys.argv) > 2 else defaults[what]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
fields = ("tbtotal", "tbatm", "tmr", "tauwet", "taudry", "tauliq")
kw = {}
if what in ("c2", "c2rt", "c2f32"):
    L = 100
    d = pkg.synth_profiles(P, L, seed=1234)
    frq, ang = pkg.HATPRO_FRQ, pkg.HATPRO_ELEV
    if what == "c2f32":
        kw["f32_lineshape"] = True
    plan = pkg.Plan("R24", L, frq, ang, ray_tracing=(what == "c2rt"), **kw)
elif what == "c3":
    L, nsub = 100, 50
    d = pkg.synth_profiles(P, L, seed=1234, clouds=True)
    frq, ang = pkg.band_frequencies(nsub), pkg.HATPRO_ELEV
    plan = pkg.Plan("R24", L, frq, ang, cloudy=True, bandpass=(14, nsub, None))
else:
    L = 150
    d = pkg.synth_profiles(P, L, seed=1234)
    frq, ang = pkg.dense_spectrum(), np.array([90.0])
    plan = pkg.Plan("R24", L, frq, ang, tau_layers=True)
    fields = ("tbtotal",)
dev = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
out = {}
ts = []
for it in range(iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = plan.forward(dev["z"], dev["p"], dev["t"], dev["rh"], dev.get("denliq"), dev.get("denice"),
                       fields=fields, out=out)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("%s P=%d L=%d F=%d A=%d: %.3f ms per forward (median of %d; all %s) launches %d checksum %.6f" % (
    what, P, L, len(frq), len(ang), float(np.median(ts[1:] or ts)), iters,
    ["%.3f" % t for t in ts], plan.launch_count(P), float(out["tbtotal"].sum())))
