#!/usr/bin/env python3
"""Time the kernels either side of the path at the reference's own size --
521 sondes x 2 crops (preprocessing4all.py:46-47) -- and at 100x that.  The recorded sondes of
tests/golden/preproc_reference.npz are replicated to make the batch.  CUDA events, median of 10.

    python tools/time_next_rows.py
"""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import mwr_fast_forward_operators_and_lbls_b200 as pkg
from mwr_fast_forward_operators_and_lbls_b200 import _lib, preproc as pp, stats
from mwr_fast_forward_operators_and_lbls_b200.forward import _ptr, _stream

g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "preproc_reference.npz"))
keys = [k for k, rej in zip(g["keys"], g["rejected"]) if not rej and not k.endswith("v")]
raws = [g[k + "_raw"] for k in keys]
N = max(r.shape[1] for r in raws)


def pad(a):
    return np.concatenate([a, np.full(N - a.size, np.nan)])


def timed(fn, reps=10):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


lib = _lib.load()
for S in (1042, 104200):
    idx = np.arange(S) % len(raws)
    H, T, P, R = (np.stack([pad(raws[i][c]) for i in idx]) for c in range(4))
    inds_one = [pp.thinning_indices(raws[i][0], raws[i][2]) for i in range(len(raws))]
    M = max(len(i) for i in inds_one)
    inds = np.zeros((S, M), dtype=np.int32)
    n_inds = np.zeros((S,), dtype=np.int32)
    for s, i in enumerate(idx):
        inds[s, :len(inds_one[i])] = inds_one[i]
        n_inds[s] = len(inds_one[i])
    dev = lambda x, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(x)).to("cuda", dt)
    h, t, p, r, di, dn = dev(H), dev(T), dev(P), dev(R), dev(inds, torch.int32), dev(n_inds, torch.int32)
    zt, tt, pt, rt = (torch.empty((S, M), dtype=torch.float64, device="cuda") for _ in range(4))
    clim = pp.midlatitude_summer()
    c = {k: dev(clim[k]) for k in ("z_km", "p", "t", "gkg", "rh")}
    K, nl = len(clim["p"]), 180
    out = [torch.empty((S, nl), dtype=torch.float64, device="cuda") for _ in range(6)]
    ok = torch.empty((S,), dtype=torch.int32, device="cuda")

    def thin():
        _lib.check(lib.mwrt_thin_mean_f64(S, N, M, _ptr(h), _ptr(t), _ptr(p), _ptr(r), _ptr(di), _ptr(dn), None,
                                          _ptr(zt), _ptr(tt), _ptr(pt), _ptr(rt), _stream()), "thin")

    def assemble():
        _lib.check(lib.mwrt_assemble_levels_f64(S, M, K, nl, float(pp.MIN_P), _ptr(zt), _ptr(tt), _ptr(pt), _ptr(rt),
                                                _ptr(dn), _ptr(c["z_km"]), _ptr(c["p"]), _ptr(c["t"]), _ptr(c["gkg"]),
                                                _ptr(c["rh"]), *[_ptr(o) for o in out], _ptr(ok), _stream()), "assemble")
    t_thin, t_asm = timed(thin), timed(assemble)
    lp, lt, lz, lrh = out[0], out[1], out[4], out[5]
    co = [torch.empty((S, nl), dtype=torch.float64, device="cuda") for _ in range(4)]
    lwp, iwp = torch.empty((S,), dtype=torch.float64, device="cuda"), torch.empty((S,), dtype=torch.float64, device="cuda")

    def cloud():
        _lib.check(lib.mwrt_cloud_features_f64(S, nl, _ptr(lp), _ptr(lt), _ptr(lz), _ptr(lrh), *[_ptr(o) for o in co],
                                               _ptr(lwp), _ptr(iwp), _stream()), "cloud")
    t_cloud = timed(cloud)
    v = torch.rand((S // 2, 14 * 10 * 2), dtype=torch.float64, device="cuda")
    t_stats = timed(lambda: stats.stats_by_channel(v, v * 1.01))
    print("S=%6d sondes x %d raw samples -> %d levels: thin_mean %.3f ms | assemble_levels %.3f ms | "
          "cloud_features %.3f ms | stats (%d x 280 columns) %.3f ms | ok %d" %
          (S, N, nl, t_thin, t_asm, t_cloud, S // 2, t_stats, int(ok.sum())), flush=True)
