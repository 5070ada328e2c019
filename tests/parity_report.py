#!/usr/bin/env python3
"""Measured parity of the CUDA path against the CPU oracle (in-repo restatement; parity with real
PyRTlib is unpinned): max relative error of the absorption coefficients and optical
depths, max |dTB|, for every model string on synthetic profiles.  Prints one JSON line per model.

    python tests/parity_report.py [P] [L] > profiles/rNN_parity.jsonl
"""
import json
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))   # repo root
import numpy as np
import mwr_fast_forward_operators_and_lbls_b200 as pkg
from oracle import pyrtlib_vec as pv

P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = int(sys.argv[2]) if len(sys.argv) > 2 else 100
d = pkg.synth_profiles(P, L, seed=4321)
e, _ = pv.vapor(d["t"], d["rh"])


def rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


for mdl in pkg.implemented_models():
    g = pkg.absorption(mdl, d["p"], d["t"], e, pkg.HATPRO_FRQ)
    awet, adry, _, _ = pv.absorption(mdl, d["p"], d["t"], e, pkg.HATPRO_FRQ)
    out = pkg.simulate(d["z"], d["p"], d["t"], d["rh"], pkg.HATPRO_FRQ, pkg.HATPRO_ELEV, mdl)
    ref = pv.simulate(d["z"], d["p"], d["t"], d["rh"], pkg.HATPRO_FRQ, pkg.HATPRO_ELEV, mdl)
    row = {"model": mdl, "profiles": P, "levels": L,
           "awet_max_rel": rel(g["awet"].cpu().numpy(), awet), "adry_max_rel": rel(g["adry"].cpu().numpy(), adry)}
    for k in ("tauwet", "taudry"):
        row[k + "_max_rel"] = rel(out[k].cpu().numpy(), ref[k])
    for k in ("tbtotal", "tbatm", "tmr"):
        row[k + "_max_abs_K"] = float(np.max(np.abs(out[k].cpu().numpy() - ref[k])))
    print(json.dumps(row), flush=True)
