#!/usr/bin/env python3
"""Development helper: MWRT_TRACE timeline of the host entry point (config 2)."""
import os, sys, time
os.environ["MWRT_TRACE"] = "1"
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
for _ in range(4):
    t0 = time.perf_counter()
    plan.forward_host(hp["z"], hp["p"], hp["t"], hp["rh"], fields=fields, out=ho)
    print("wall %.3f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
