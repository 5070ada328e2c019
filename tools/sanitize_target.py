#!/usr/bin/env python3
"""compute-sanitizer target: every kernel of libmwrt.so once, at the smallest sizes that still take
every code path (clear sky, cloudy + fused band-pass, unfused band-pass fallback, ray tracing incl. a
ducted profile, fp32 line-shape mode, tmrcld, upwelling, host entry point with a ragged last chunk,
per-layer optical depth, NaN gate, post-processing kernels).  Run as

    compute-sanitizer --tool memcheck  python tools/sanitize_target.py
    compute-sanitizer --tool racecheck python tools/sanitize_target.py

(one tool per gpurun call, see /opt/skills/guides/B200_PROFILING.md).  Prints a checksum per case so a
silent skip is visible in the log.
"""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import mwr_fast_forward_operators_and_lbls_b200 as pkg
from mwr_fast_forward_operators_and_lbls_b200 import stats

FRQ, ANG = pkg.HATPRO_FRQ, pkg.HATPRO_ELEV


def show(name, t):
    t = t if isinstance(t, torch.Tensor) else torch.as_tensor(t)
    print("%-28s checksum %.9e" % (name, float(torch.nan_to_num(t.double(), nan=0.0).sum())), flush=True)


d = pkg.synth_profiles(37, 33, seed=5, clouds=True)          # odd sizes: ragged tiles everywhere
d["t"][3, 4] = np.nan
out = pkg.simulate(d["z"], d["p"], d["t"], d["rh"], FRQ, ANG, "R24")
show("clear R24", out["tbtotal"])
out = pkg.simulate(d["z"], d["p"], d["t"], d["rh"], FRQ, ANG, "R98", from_sat=True)
show("upwelling R98", out["tbtotal"])
out = pkg.simulate(d["z"], d["p"], d["t"], d["rh"], FRQ, ANG[:3], "R22SD", denliq=d["denliq"], denice=d["denice"])
show("cloudy R22SD", out["tauliq"])
out = pkg.simulate(d["z"], d["p"], d["t"], d["rh"], FRQ, ANG, "R24", f32_lineshape=True)
show("fp32 line shape", out["tbtotal"])
frq = pkg.band_frequencies(5)
out = pkg.simulate(d["z"], d["p"], d["t"], d["rh"], frq, ANG, "R24", denliq=d["denliq"], denice=d["denice"],
                   bandpass=(14, 5, None))
show("band-pass fused 14x5", out["tbtotal"])
frq = pkg.band_frequencies(400, FRQ[:1], np.array([0.23]))
out = pkg.simulate(d["z"][:5], d["p"][:5], d["t"][:5], d["rh"][:5], frq, ANG[:2], "R17", bandpass=(1, 400, None))
show("band-pass fallback 1x400", out["tbtotal"])
d["rh"][2, :6] = 1.0
d["rh"][2, 6:] = 0.01
out = pkg.simulate(d["z"], d["p"], d["t"], d["rh"], FRQ, np.concatenate([ANG, [0.3, 89.7]]), "R20", ray_tracing=True)
show("ray tracing (+duct)", out["tbtotal"])
show("  status", out["status"])
out = pkg.simulate(d["z"], d["p"], d["t"], d["rh"], np.linspace(20, 60, 130), [90.0], "R24", tau_layers=True)
show("130 freqs + tau_layer", out["tau_layer"])
i = int(np.argmax(d["denliq"].max(axis=1)))
wet = np.where(d["denliq"][i] > 0)[0]
rte = pkg.TbCloudRTE(d["z"][i], d["p"][i], d["t"][i], d["rh"][i], FRQ, [90.0, 19.2], from_sat=False)
rte.init_absmdl("R24")
rte.init_cloudy(np.array([[d["z"][i][wet[0]]], [d["z"][i][wet[-1]]]]), d["denice"][i], d["denliq"][i])
show("class API tmrcld", rte.execute()["tmrcld"].values)
big = pkg.synth_profiles(2300, 20, seed=6)
plan = pkg.Plan("R24", 20, FRQ, ANG[:4])
res = plan.forward_host(big["z"], big["p"], big["t"], big["rh"])
show("host entry 2300 profiles", res["tbtotal"])
e, rho = pkg.vapor(d["t"], d["rh"])
show("vapor", e)
a = pkg.absorption("R24", d["p"], d["t"], e, FRQ, d["denliq"], d["denice"])
show("stage absorption aliq", a["aliq"])
ds, st = pkg.slant_path(d["z"], d["p"], d["t"], e, ANG, True)
show("stage slant path", ds)
r = pkg.rte(d["t"], a["awet"], a["adry"], ds, FRQ, aliq=a["aliq"], aice=a["aice"], tau_layer_angle=1)
show("stage rte (arrays)", r["tbtotal"])
show("stage bandpass", pkg.bandpass(r["tbtotal"], 7, 2))
v = torch.rand((11, 140), dtype=torch.float64, device="cuda")
s = stats.stats_by_channel(v, v * 1.01)
show("stats", s[2])
torch.cuda.synchronize()
print("sanitize target: done", flush=True)
