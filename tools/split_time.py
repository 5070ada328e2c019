#!/usr/bin/env python3
"""Development helper: effect of mwrt_plan_set_split on the device and host entry points (config 2)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import mwr_fast_forward_operators_and_lbls_b200 as pkg

P, L = 10000, 100
d = pkg.synth_profiles(P, L, seed=1234)
dev = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
hp = pkg.pinned_fields(list(d.keys()), d["z"].shape)
for k in d: hp[k][...] = d[k]
fields = ("tbtotal", "tbatm", "tmr", "tauwet", "taudry", "tauliq")
ho = pkg.pinned_fields(fields, (P, 14, 10))
ho["status"] = torch.empty((P,), dtype=torch.int32).pin_memory().numpy()
ref = None
for split in (1, 2, 3, 4, 6, 8):
    plan = pkg.Plan("R24", L, pkg.HATPRO_FRQ, pkg.HATPRO_ELEV)
    plan.set_split(split)
    out = {}
    for _ in range(3):
        out = plan.forward(dev["z"], dev["p"], dev["t"], dev["rh"], fields=fields, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.forward(dev["z"], dev["p"], dev["t"], dev["rh"], fields=fields, out=out); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    tb = out["tbtotal"].cpu().numpy()
    if ref is None:
        ref = tb
    same = np.array_equal(ref, tb)
    for _ in range(3):
        plan.forward_host(hp["z"], hp["p"], hp["t"], hp["rh"], fields=fields, out=ho)
    hs = []
    for _ in range(20):
        t0 = time.perf_counter(); plan.forward_host(hp["z"], hp["p"], hp["t"], hp["rh"], fields=fields, out=ho)
        hs.append((time.perf_counter() - t0) * 1e3)
    print("split %d: device %.3f ms (min %.3f) identical=%s | host e2e %.3f ms (min %.3f) host==dev %s"
          % (split, np.median(ts), min(ts), same, np.median(hs), min(hs), np.array_equal(ho["tbtotal"], ref)))
