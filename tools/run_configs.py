#!/usr/bin/env python3
"""Measure the BASELINE.json configurations that are not the bench line (configs[2..4]) on ONE GPU
(the per-GPU share of the sharded configurations) and print one JSON object per configuration.
Timing: CUDA events around Plan.forward with device-resident inputs, 2 warm-ups, median of 3-5.

    python tools/run_configs.py [3] [4] [5] [--scale 1.0]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import mwr_fast_forward_operators_and_lbls_b200 as pkg  # noqa: E402
from mwr_fast_forward_operators_and_lbls_b200 import roofline  # noqa: E402


def timed(fn, reps):
    fn()
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), ts


def run(cfg, scale):
    peak, _ = pkg.measure_dfma_peak()
    if cfg == 3:
        # 10^5 profiles with liquid cloud, 14 channels x 50 sub-frequencies, band-pass -> 14; 1/8 share
        P, L, nsub = int(12500 * scale), 100, 50
        d = pkg.synth_profiles(P, L, seed=1234, clouds=True)
        frq, ang = pkg.band_frequencies(nsub), pkg.HATPRO_ELEV
        plan = pkg.Plan("R24", L, frq, ang, cloudy=True, bandpass=(14, nsub, None))
        fields = ("tbtotal", "tbatm", "tmr", "tauwet", "taudry", "tauliq")
        name = "configs[2]: 1/8 share (12500 of 10^5 profiles), liquid cloud, 700 sub-frequencies -> 14 channels"
        kw = dict(cloudy=True)
    elif cfg == 4:
        # dense spectrum 20-60 GHz @ 10 MHz (4001) x 10^4 profiles x 150 levels, tau + TB; 1 GPU
        P, L = int(10000 * scale), 150
        d = pkg.synth_profiles(P, L, seed=1234)
        frq, ang = pkg.dense_spectrum(), np.array([90.0])
        plan = pkg.Plan("R24", L, frq, ang, tau_layers=True)
        fields = ("tbtotal",)
        name = "configs[3]: dense spectrum 4001 freqs x %d profiles x 150 levels, per-layer tau (P,F,L) + zenith TB" % P
        kw = dict(tau_layers=True)
    else:
        # ICON-D2-like: 5x10^5 columns x 65 levels; 1/8 share
        P, L = int(62500 * scale), 65
        d = pkg.synth_profiles(P, L, seed=1234, top_km=22.0, exponent=2.5, lowest_layer_km=0.02)
        frq, ang = pkg.HATPRO_FRQ, pkg.HATPRO_ELEV
        plan = pkg.Plan("R24", L, frq, ang)
        fields = ("tbtotal", "tbatm", "tmr", "tauwet", "taudry", "tauliq")
        name = "configs[4]: 1/8 share (62500 of 5x10^5 ICON-D2-like columns), 65 levels"
        kw = {}
    dev = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    F, A = len(frq), len(ang)
    out = {}

    def fn():
        nonlocal out
        out = plan.forward(dev["z"], dev["p"], dev["t"], dev["rh"], dev.get("denliq"), dev.get("denice"),
                           fields=fields, out=out)
    t0 = time.time()
    ms, ts = timed(fn, 3 if cfg == 4 else 5)
    w = roofline.work(P, L, F, A, cloudy=kw.get("cloudy", False), tau_layers=kw.get("tau_layers", False))
    ntb = P * plan.Fo * A
    res = {"config": name, "P": P, "L": L, "F": F, "A": A, "ms": ms, "all_ms": ts,
           "tb_per_s": ntb / ms * 1e3, "mono_tb_per_s": P * F * A / ms * 1e3,
           "algorithmic_tflops": w["flops_total"] / ms / 1e9, "dfma_peak_tflops": peak / 1e12,
           "frac_of_dfma_peak": w["flops_total"] / ms / 1e9 / (peak / 1e12),
           "out_gbytes": (w["bytes_out"]) / 1e9, "out_gbs": w["bytes_out"] / ms / 1e6,
           "finite": bool(torch.isfinite(out["tbtotal"]).all().item()),
           "wall_s": time.time() - t0}
    return res


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    scale = 1.0
    if "--scale" in sys.argv:
        scale = float(sys.argv[sys.argv.index("--scale") + 1])
        args = [a for a in args if a != str(scale)]
    for c in [int(a) for a in args] or [3, 4, 5]:
        print(json.dumps(run(c, scale)), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
