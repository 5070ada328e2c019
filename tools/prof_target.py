#!/usr/bin/env python3
"""ncu target: config 2 (10^4 profiles x 100 levels x 14 channels x 10 elevations, R24) through the
plan, one kernel-1 and one kernel-3 launch per forward (split 1) so a capture sees the full grids.

    python tools/prof_target.py [P] [L] [iters]
"""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import mwr_fast_forward_operators_and_lbls_b200 as pkg

P = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 100
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
d = pkg.synth_profiles(P, L, seed=1234)
dev = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
plan = pkg.Plan("R24", L, pkg.HATPRO_FRQ, pkg.HATPRO_ELEV)
plan.set_split(1)
out = {}
for _ in range(iters):
    out = plan.forward(dev["z"], dev["p"], dev["t"], dev["rh"], out=out)
torch.cuda.synchronize()
print("checksum", float(out["tbtotal"].sum()))
