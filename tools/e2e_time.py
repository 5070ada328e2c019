#!/usr/bin/env python3
"""Development helper: host entry point timing (config 2), `python tools/e2e_time.py [split]`."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import mwr_fast_forward_operators_and_lbls_b200 as pkg
P, L = 10000, 100
d = pkg.synth_profiles(P, L, seed=1234)
hp = pkg.pinned_fields(list(d.keys()), d["z"].shape)
for k in d: hp[k][...] = d[k]
fields = ("tbtotal", "tbatm", "tmr", "tauwet", "taudry", "tauliq")
ho = pkg.pinned_fields(fields, (P, 14, 10))
ho["status"] = torch.empty((P,), dtype=torch.int32).pin_memory().numpy()
plan = pkg.Plan("R24", L, pkg.HATPRO_FRQ, pkg.HATPRO_ELEV)
if len(sys.argv) > 1: plan.set_split(int(sys.argv[1]))
for _ in range(5): plan.forward_host(hp["z"], hp["p"], hp["t"], hp["rh"], fields=fields, out=ho)
hs = []
for _ in range(30):
    t0 = time.perf_counter(); plan.forward_host(hp["z"], hp["p"], hp["t"], hp["rh"], fields=fields, out=ho)
    hs.append((time.perf_counter() - t0) * 1e3)
print("streams=%s split=%s: host e2e %.3f ms (min %.3f) -> %.3e TB/s" % (os.environ.get("MWRT_HOST_STREAMS", "2"), sys.argv[1] if len(sys.argv) > 1 else "-", np.median(hs), min(hs), P * 140 / np.median(hs) * 1e3))
