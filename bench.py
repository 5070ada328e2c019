#!/usr/bin/env python3
"""Headline benchmark: profile*channel*angle brightness temperatures per second (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)
    python bench.py --config 3|4|5                            # another BASELINE config as the headline line

Workload (config.workload): BASELINE.json configs[1] -- 10^4 synthetic radiosonde-shaped profiles
(100 levels) x 14 HATPRO channels x 10 elevations, clear sky, model R24 (the reference's analysis
baseline, PyRTlib_processing.py:128-135), PER GPU (weak scaling: profiles shard by rank, the only
exchange is the final gather of the results).  One "step" = one pass of the whole path
(vapour pressure -> absorption -> slant path -> RTE) over that batch.

  value  device-resident inputs; ONE CUDA-event pair around the K steps (L2 flushes and the tail of the
         last gather inside it), max over ranks -- the same clock for every N
  e2e    the same batch through the host entry point (mwrt_plan_forward_host): pinned host buffers in,
         host buffers out, H2D and D2H copies inside the timed region
  roofline  dominant kernel = k_absorption, timed alone with CUDA events; fp64-pipe bound, so the
         denominator is a DFMA-stream peak measured in the same run (MEASURED_PEAKS.json has no fp64
         figure); HBM numbers are given beside it
  cpu_baseline  oracle/cpu_baseline.py (PyRTlib-shaped port, all host cores) on a bounded sample; the
         same leg checks the GPU against the oracle on 256 seeded profiles -> max_abs_dtb_vs_oracle_K
  extra.configs  BASELINE.json configs[2..4] at FULL size, split over the N ranks (strong scaling) and
         gathered: 10^5 cloudy profiles with 700 sub-frequencies -> 14 channels; the 4001-frequency
         dense spectrum with per-layer optical depth; 5x10^5 ICON-D2-like columns
  extra.f32_lineshape  the fp32 line-shape mode, reported separately (north_star: <= 0.01 K)
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "profile*channel*angle TBs/sec"
UNIT = "TB/s (brightness temperatures per second)"
P_PER_GPU, LEVELS, MODEL = 10000, 100, "R24"
WORKLOAD = ("BASELINE.json configs[1]: 10^4 synthetic radiosonde-shaped profiles (100 levels) x 14 "
            "HATPRO channels x 10 elevations, clear sky, R24, per GPU")
SIX = ("tbtotal", "tbatm", "tmr", "tauwet", "taudry", "tauliq")   # north_star's six outputs


_JSON_OUT = None


def _reserve_stdout():
    """Keep the real stdout for the ONE JSON line and point fd 1 at stderr for everything else, so that
    library chatter written to the C-level stdout (NCCL prints its version banner there) cannot end up
    beside the JSON line."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def cpu_arm(steps, warmup, budget_s, parity_out=""):
    cmd = [sys.executable, "-m", "oracle.cpu_baseline", "--levels", str(LEVELS), "--model", MODEL,
           "--steps", str(steps), "--warmup", str(warmup), "--budget-s", str(budget_s)]
    if parity_out:
        cmd += ["--parity-out", parity_out]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("cpu baseline failed: " + r.stderr[-2000:])
    return json.loads(r.stdout.strip().splitlines()[-1])


def run_reference(args):
    """The reference's own CPU implementation of the path (PyRTlib-shaped oracle port; real pyrtlib
    if the box has it), all host cores, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_arm(args.steps, args.warmup, budget_s=120.0)
    ms = 1e3 * sum(res["step_s"]) / len(res["step_s"])
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": WORKLOAD, "model": MODEL, "levels": LEVELS,
                                            "channels": 14, "elevations": 10},
            "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"],
                             "kind": res["kind"], "sample": res["sample"]},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples = []
        self.proc = None
        self.index = index

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append((time.time(), [c.strip() for c in line.split(",")]))
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, t0, t1):
        import statistics
        win = [s for (t, s) in self.samples if t0 <= t <= t1] or [s for (_, s) in self.samples]
        if not win:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in win for n, v in zip(names, s[3:7]) if v.lower().startswith("active")})

        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = [num(s[0]) for s in win if num(s[0]) is not None]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": num(win[0][1]),
                "power_w_max": max([num(s[2]) or 0.0 for s in win]), "reasons": reasons,
                "samples": len(win)}


# --------------------------------------------------------------------------------------------------
# BASELINE.json configs[2..4] at full size, split over the ranks
# --------------------------------------------------------------------------------------------------
CONFIG_NAMES = {
    2: WORKLOAD,
    3: "BASELINE.json configs[2]: 10^5 profiles with liquid-cloud layers, 14 channels x 50 sub-frequencies "
       "(700 monochromatic) band-pass averaged to 14, 10 elevations, R24, split over the ranks",
    4: "BASELINE.json configs[3]: dense spectrum 20-60 GHz at 10 MHz (4001 frequencies) x 10^4 profiles x 150 "
       "levels, per-layer optical depth (P,F,L) + zenith TB, R24, split over the ranks",
    5: "BASELINE.json configs[4]: ICON-D2-like domain, 5x10^5 columns x 65 levels x 14 channels x 10 "
       "elevations, R24, split over the ranks",
}


def build_config(pkg, cfg, rank, world, scale=1.0):
    """Plan + this rank's shard of configuration `cfg` (profiles split evenly; seed = 1234 + rank)."""
    import numpy as np
    from mwr_fast_forward_operators_and_lbls_b200.distributed import shard_range
    if cfg == 3:
        Ptot, L, nsub = int(100000 * scale), 100, 50
        a, b = shard_range(Ptot, rank, world)
        d = pkg.synth_profiles(b - a, L, seed=1234 + rank, clouds=True)
        frq, ang = pkg.band_frequencies(nsub), pkg.HATPRO_ELEV
        plan = pkg.Plan(MODEL, L, frq, ang, cloudy=True, bandpass=(14, nsub, None))
        fields, kw = SIX, dict(cloudy=True)
    elif cfg == 4:
        Ptot, L = int(10000 * scale), 150
        a, b = shard_range(Ptot, rank, world)
        d = pkg.synth_profiles(b - a, L, seed=1234 + rank)
        frq, ang = pkg.dense_spectrum(), np.array([90.0])
        plan = pkg.Plan(MODEL, L, frq, ang, tau_layers=True)
        fields, kw = ("tbtotal",), dict(tau_layers=True)
    elif cfg == 5:
        Ptot, L = int(500000 * scale), 65
        a, b = shard_range(Ptot, rank, world)
        d = pkg.synth_profiles(b - a, L, seed=1234 + rank, top_km=22.0, exponent=2.5, lowest_layer_km=0.02)
        frq, ang = pkg.HATPRO_FRQ, pkg.HATPRO_ELEV
        plan = pkg.Plan(MODEL, L, frq, ang)
        fields, kw = SIX, {}
    else:
        raise ValueError(cfg)
    return dict(plan=plan, d=d, Ptot=Ptot, P=b - a, L=L, F=len(frq), A=len(ang), fields=fields, kw=kw)


def run_config(pkg, cfg, rank, world, reps=3, scale=1.0, gather="push", pieces=4):
    """One BASELINE configuration at full size over `world` ranks: every rank computes its shard and the
    (P,Fo,A) results end up on EVERY rank, inside the timing -- by default pushed piece by piece while
    the next piece computes (distributed.PushGatherPass; checked once against an NCCL all-gather of the
    same fields), with gather="nccl" by one all_gather_into_tensor after the kernels.  Config 4's 48 GB
    per-layer optical depth stays sharded (each rank keeps its block).
    Time = CUDA events around a pass, median of `reps`, max over ranks."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from mwr_fast_forward_operators_and_lbls_b200 import roofline
    from mwr_fast_forward_operators_and_lbls_b200.distributed import PushGatherPass
    c = build_config(pkg, cfg, rank, world, scale)
    plan, d, fields = c["plan"], c["d"], c["fields"]
    P, Ptot, L, F, A = c["P"], c["Ptot"], c["L"], c["F"], c["A"]
    dev = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    nf, Fo = len(fields), plan.Fo
    pmax = -(-Ptot // world)
    local_out = {}
    if c["kw"].get("tau_layers"):
        local_out["tau_layer"] = torch.empty((P, F, L), dtype=torch.float64, device="cuda")
    pg, how = None, "single GPU"
    if world > 1 and gather == "push" and Ptot % world == 0:
        try:
            pg = PushGatherPass(plan, P, fields, pieces=pieces)
            how = ("push gather in %d pieces (%s), overlapped with the kernels of the next piece; all ranks hold "
                   "the (P,Fo,A) fields" % (pg.pieces, pg.transport))
        except Exception as exc:      # no symmetric memory: plain NCCL
            if rank == 0:
                print("push gather unavailable for config %d (%r); using NCCL" % (cfg, exc), file=sys.stderr)
            pg = None
    out = {k: torch.empty((P, Fo, A), dtype=torch.float64, device="cuda") for k in fields}
    out["status"] = torch.empty((P,), dtype=torch.int32, device="cuda")
    out.update(local_out)
    stacked = full = None
    if world > 1:
        stacked = torch.zeros((pmax, nf, Fo, A), dtype=torch.float64, device="cuda")   # padded to the largest shard
        full = torch.empty((world * pmax, nf, Fo, A), dtype=torch.float64, device="cuda")
        if pg is None:
            how = "one all_gather_into_tensor of the stacked fields after the kernels; all ranks hold the (P,Fo,A) fields"

    def nccl_pass():
        plan.forward(dev["z"], dev["p"], dev["t"], dev["rh"], dev.get("denliq"), dev.get("denice"),
                     fields=fields, out=out)
        if world > 1:
            for i, k in enumerate(fields):
                stacked[:P, i].copy_(out[k])
            dist.all_gather_into_tensor(full, stacked)
    ts, tc = [], []
    for it in range(1 + reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, em, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        if pg is not None:
            pg.run(dev["z"], dev["p"], dev["t"], dev["rh"], dev.get("denliq"), dev.get("denice"), local_out=local_out)
            em.record()
            pg.finish()
        else:
            plan.forward(dev["z"], dev["p"], dev["t"], dev["rh"], dev.get("denliq"), dev.get("denice"),
                         fields=fields, out=out)
            em.record()
            if world > 1:
                for i, k in enumerate(fields):
                    stacked[:P, i].copy_(out[k])
                dist.all_gather_into_tensor(full, stacked)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1), e0.elapsed_time(em)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it > 0:
            ts.append(float(t[0].item()))
            tc.append(float(t[1].item()))
    ms = float(np.median(ts))
    identical = None
    if pg is not None:      # the pushed result against NCCL's, once, outside the timing
        nccl_pass()
        torch.cuda.synchronize()
        same = True
        for i, k in enumerate(fields):
            ref = full[:, i].reshape(world, P, Fo, A)
            got = pg.field(k)                                  # (world, pieces, Pp, Fo, A)
            got = got.reshape(world, pg.pieces * pg.Pp, Fo, A)[:, :P]
            same = same and bool(torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(ref, nan=-7.0)))
        flag = torch.tensor([1 if same else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        identical = bool(flag.item())
        tb_ref = pg.field("tbtotal")
        finite = bool(torch.isfinite(tb_ref.reshape(world, -1, Fo, A)[:, :P]).all().item())
    else:
        finite = bool(torch.isfinite(out["tbtotal"]).all().item())
    w = roofline.work(Ptot, L, F, A, cloudy=c["kw"].get("cloudy", False), tau_layers=c["kw"].get("tau_layers", False))
    res = {"config": CONFIG_NAMES[cfg], "profiles_total": Ptot, "profiles_per_rank": P, "levels": L,
           "frequencies": F, "channels_out": Fo, "elevations": A, "n_gpus": world, "ms_per_pass": ms,
           "ms_kernels_only": float(np.median(tc)), "gathered_gbytes_per_rank": 8.0 * nf * Ptot * Fo * A / 1e9,
           "all_ms": ts, "tb_per_s": Ptot * Fo * A / ms * 1e3, "mono_tb_per_s": Ptot * F * A / ms * 1e3,
           "algorithmic_tflops": w["flops_total"] / ms / 1e9,
           "algorithmic_tflops_per_gpu": w["flops_total"] / ms / 1e9 / world,
           "gathered": how + ("; tau_layer stays sharded" if cfg == 4 and world > 1 else ""),
           "gather_identical_to_nccl": identical,
           "launches_per_pass_per_rank": plan.launch_count(P), "finite": finite}
    if cfg == 4:
        res["tau_layer_gbytes"] = 8.0 * Ptot * F * L / 1e9
        res["tau_layer_gbs"] = 8.0 * Ptot * F * L / ms / 1e6
    del plan, dev, out, stacked, full, pg
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip extra.configs (BASELINE configs 3/4/5 at full size) and extra.f32_lineshape")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="2 (default): the headline workload, BASELINE.json configs[1]; 3/4/5: print the line "
                         "of that configuration instead (full size, split over the ranks)")
    ap.add_argument("--scale", type=float, default=1.0, help="--config 3/4/5 and extras: fraction of the full size")
    ap.add_argument("--model", default=MODEL)
    ap.add_argument("--gather", default="push", choices=["push", "ce", "fused", "nccl"],
                    help="N>1: 'push' = a small kernel stores each rank's results to the NVLink multicast "
                         "address of a symmetric buffer while the next step computes (per-peer stores or copy "
                         "engines where multicast is missing); 'ce' = one copy-engine memcpy per peer; 'fused' = "
                         "kernel 3 itself stores to every rank's buffer; 'nccl' = all_gather_into_tensor "
                         "overlapped with the next step")
    ap.add_argument("--split", type=int, default=1,
                    help="mwrt_plan_set_split: pieces of the device-entry batch (kernel 3 of piece c on a side "
                         "stream while kernel 1 of piece c+1 runs)")
    args = ap.parse_args()
    _reserve_stdout()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    import mwr_fast_forward_operators_and_lbls_b200 as pkg
    from mwr_fast_forward_operators_and_lbls_b200 import _lib, hostbind, roofline
    from mwr_fast_forward_operators_and_lbls_b200.distributed import (FusedGatherForward, PipelinedForward,
                                                                     PushGather)
    from mwr_fast_forward_operators_and_lbls_b200.forward import _ptr, _stream

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    # the host entry point's pinned buffers are first-touched on the GPU's own NUMA node (hostbind.py);
    # MWRT_NO_BIND=1 keeps the inherited affinity (for the with/without comparison of tools/e2e_diag.py)
    orig_affinity = os.sched_getaffinity(0)
    numa = ({"bound": False, "why": "MWRT_NO_BIND"} if os.environ.get("MWRT_NO_BIND")
            else hostbind.bind_to_gpu_node(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N = world

    if args.config != 2:       # another configuration as the line (builder runs; the driver uses the default)
        r = run_config(pkg, args.config, rank, world, reps=max(3, min(args.steps, 10)), scale=args.scale,
                       gather="nccl" if args.gather == "nccl" else "push")
        if rank == 0:
            _emit({"metric": METRIC, "value": r["tb_per_s"], "unit": UNIT, "n_gpus": N, "steps": len(r["all_ms"]),
                   "warmup": 1, "ms_per_step": r["ms_per_pass"], "higher_is_better": True, "scaling": "strong",
                   "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                   "config": {"workload": r["config"], "scale": args.scale}, "detail": r,
                   "gpu_launches": len(r["all_ms"]) * r["launches_per_pass_per_rank"] * N})
        if N > 1:
            dist.destroy_process_group()
        return

    P, L = P_PER_GPU, LEVELS
    frq, ang = pkg.HATPRO_FRQ, pkg.HATPRO_ELEV
    F, A = len(frq), len(ang)

    d = pkg.synth_profiles(P, L, seed=1234 + rank)
    keys = ("z", "p", "t", "rh")
    dev = {k: torch.from_numpy(d[k]).cuda() for k in keys}
    plan = pkg.Plan(args.model, L, frq, ang)
    if args.split > 1:
        plan.set_split(args.split)
    fields = SIX
    out = {k: torch.empty((P, F, A), dtype=torch.float64, device="cuda") for k in fields}
    out["status"] = torch.empty((P,), dtype=torch.int32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # > 126 MB L2

    # N>1: the steps are a stream of batches; the gather of step s (the fields that carry information +
    # the status words) travels while step s+1 computes (double-buffered).
    ovl = None
    gather_desc = "none"
    if N > 1:
        if args.gather in ("push", "ce"):
            try:
                ovl = PushGather(plan, P, fields, transport="auto" if args.gather == "push" else "ce")
                how = {"multicast": "a %d-CTA kernel stores the slot to the buffer's NVLink multicast address "
                                    "(multimem.st; NVSwitch replicates it to all ranks, a rank sends once)" % ovl.nblocks,
                       "peers": "a %d-CTA kernel stores the slot to every peer-mapped copy" % ovl.nblocks,
                       "ce": "one cudaMemcpyAsync per peer on side streams (copy engines)"}[ovl.transport]
                gather_desc = ("push gather: kernel 3 writes its %d fields (%s; tauliq is identically zero in clear "
                               "sky and is rebuilt from the gathered status words) into the rank's slot of a "
                               "symmetric buffer; %s while the next step computes; one symmetric-memory barrier "
                               "per step; no collective call" % (len(ovl.moved), ",".join(ovl.moved), how))
            except Exception as exc:          # symmetric memory not available: plain NCCL
                if rank == 0:
                    print("push gather unavailable (%r); using NCCL" % (exc,), file=sys.stderr)
                ovl = None
        elif args.gather == "fused":
            try:
                ovl = FusedGatherForward(plan, P, fields)
                gather_desc = ("fused into kernel 3: every rank's RTE kernel stores its results into the "
                               "gathered buffer of all ranks over NVLink (%s), one symmetric-memory "
                               "barrier per step; no collective kernel" %
                               ("multimem.st multicast through NVSwitch" if ovl.mode == "multicast"
                                else "one store per peer"))
            except Exception as exc:
                if rank == 0:
                    print("fused gather unavailable (%r); using NCCL" % (exc,), file=sys.stderr)
                ovl = None
        if ovl is None:
            ovl = PipelinedForward(plan, P, fields)
            gather_desc = ("one all_gather_into_tensor of the six stacked outputs per step, overlapped "
                           "with the next step's kernels (double-buffered)")

    def step():
        if N > 1:
            return ovl.run(dev["z"], dev["p"], dev["t"], dev["rh"])
        return plan.forward(dev["z"], dev["p"], dev["t"], dev["rh"], fields=fields, out=out)

    def barrier():
        if N > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak_flops, _ = pkg.measure_dfma_peak()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        flush.zero_()
        step()
    if N > 1:
        ovl.finish()
    barrier()
    # ONE event pair around the whole timed region -- K steps, the L2 flush before each of them and the
    # tail of the last gather -- for every N, so that nothing runs un-timed between steps
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    t_begin.record()
    for _ in range(args.steps):
        flush.zero_()          # L2 flush between timed steps (inside the timed region)
        res = step()
    if N > 1:
        ovl.finish()           # the last step's gather must have landed before the clock stops
    t_end.record()
    barrier()
    t_wall1 = time.time()
    total_ms = torch.tensor([t_begin.elapsed_time(t_end)], dtype=torch.float64, device="cuda")
    if N > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = N * P * F * A * args.steps / (total_ms * 1e-3)
    if N > 1:
        checksum = float(ovl.field("tbtotal", res).sum().item())
    else:
        checksum = float(res["tbtotal"].sum().item())
    # what the flushes cost inside the timed region (reported, not subtracted)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(20):
        flush.zero_()
    f1.record()
    torch.cuda.synchronize()
    flush_ms = f0.elapsed_time(f1) / 20

    # ---- end to end through the host entry point (pinned host memory, copies inside the timing)
    hin = pkg.pinned_fields(keys, d[keys[0]].shape)     # one pinned block per direction: the library
    for k in keys:                                      # moves each chunk with one 2-D copy
        hin[k][...] = d[k]
    hout = pkg.pinned_fields(fields, (P, F, A))
    hout["status"] = torch.empty((P,), dtype=torch.int32).pin_memory().numpy()
    e2e_steps = max(3, min(args.steps, 50))
    for _ in range(3):
        plan.forward_host(hin["z"], hin["p"], hin["t"], hin["rh"], fields=fields, out=hout)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        plan.forward_host(hin["z"], hin["p"], hin["t"], hin["rh"], fields=fields, out=hout)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if N > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = N * P * F * A * e2e_steps / float(e2e_s.item())
    h2d = 4 * 8 * P * L
    # clear sky: tauliq is identically zero and is filled on the host, not moved over PCIe
    moved = [f for f in fields if f not in ("tauliq", "tauice")]
    d2h = len(moved) * 8 * P * F * A + 4 * P
    # the reference's own driver reads only `tbtotal` (PyRTlib_processing.py:127): same call, one field
    hout1 = {"tbtotal": hout["tbtotal"], "status": hout["status"]}
    for _ in range(3):
        plan.forward_host(hin["z"], hin["p"], hin["t"], hin["rh"], fields=("tbtotal",), out=hout1)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        plan.forward_host(hin["z"], hin["p"], hin["t"], hin["rh"], fields=("tbtotal",), out=hout1)
    barrier()
    e2e1_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if N > 1:
        dist.all_reduce(e2e1_s, op=dist.ReduceOp.MAX)
    e2e1_value = N * P * F * A * e2e_steps / float(e2e1_s.item())

    # ---- dominant kernel alone (k_absorption) for the roofline, fp64 and fp32 line-shape mode
    lib = _lib.load()
    m = pkg.get_model(args.model)
    e_t = pkg.vapor(dev["t"], dev["rh"])[0]
    f_t = torch.from_numpy(frq).cuda()
    awet = torch.empty((P, F, L), dtype=torch.float64, device="cuda")
    adry = torch.empty_like(awet)

    def time_abs(pl):
        # kernel 1 alone through the plan's stage call (mwrt_plan_absorption): the same launch the forward
        # pass makes, per-block tables cached in the plan
        def k_abs():
            pl.absorption(dev["p"], dev["t"], e_t, out=(awet, adry))
        for _ in range(3):
            k_abs()
        torch.cuda.synchronize()
        kt = []
        for _ in range(max(5, min(args.steps, 30))):
            flush.zero_()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            k_abs()
            a1.record()
            torch.cuda.synchronize()
            kt.append(a0.elapsed_time(a1))
        return sum(kt) / len(kt)
    k_ms = time_abs(plan)
    sampler.stop()
    desc = pkg.load_descriptor(args.model)
    w = roofline.work(P, L, F, A, n_h2o=len(desc["h2o"]["fl"]), n_o2=len(desc["o2"]["f"]))
    achieved = w["flops_k_abs"] / (k_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    traffic, traffic_src = None, None   # measured DRAM bytes per launch from the committed ncu --set full capture
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            if args.model == MODEL:
                traffic = json.load(open(os.path.join(ROOT, "profiles", name)))["k_absorption"]["traffic_bytes"]
                traffic_src = "profiles/" + name
                break
        except Exception:
            pass
    step_ms = total_ms / args.steps
    roof = {"bound": "fp64", "kernel": "k_absorption", "achieved": achieved,
            "peak": peak_flops / 1e12, "unit": "TFLOP/s", "frac": achieved / (peak_flops / 1e12),
            "peak_source": "self-measured DFMA stream in this run (nominal 37.2); MEASURED_PEAKS.json "
                           "has no fp64 figure",
            "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram read+write, %s)" % traffic_src,
            "kernel_ms": k_ms,
            "kernel_share_of_step": k_ms / (step_ms - flush_ms),
            "algorithmic_flops_per_launch": w["flops_k_abs"],
            "hbm": {"algorithmic_bytes_per_launch": w["bytes_k_abs"],
                    "achieved_gbs": w["bytes_k_abs"] / (k_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
            "whole_step_tflops": w["flops_total"] / ((step_ms - flush_ms) * 1e-3) / 1e12}

    # ---- extras: fp32 line-shape mode (reported separately) and BASELINE configs 3/4/5 at full size
    extra = {}
    if not args.no_extras:
        plan32 = pkg.Plan(args.model, L, frq, ang, f32_lineshape=True)
        k32_ms = time_abs(plan32)
        out32 = {k: torch.empty((P, F, A), dtype=torch.float64, device="cuda") for k in fields}
        out32["status"] = torch.empty((P,), dtype=torch.int32, device="cuda")
        o64 = plan.forward(dev["z"], dev["p"], dev["t"], dev["rh"], fields=fields, out=out)
        for _ in range(3):
            plan32.forward(dev["z"], dev["p"], dev["t"], dev["rh"], fields=fields, out=out32)
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n32 = max(5, min(args.steps, 30))
        g0.record()
        for _ in range(n32):
            flush.zero_()
            plan32.forward(dev["z"], dev["p"], dev["t"], dev["rh"], fields=fields, out=out32)
        g1.record()
        torch.cuda.synchronize()
        ms32 = g0.elapsed_time(g1) / n32
        dtb = float((out32["tbtotal"] - o64["tbtotal"]).abs().max().item())
        dtau = float(((out32["taudry"] + out32["tauwet"]) / (o64["taudry"] + o64["tauwet"]) - 1).abs().max().item())
        extra["f32_lineshape"] = {
            "what": "fp32 line-shape mode (mwrt_plan_set_precision / mwrt_absorption_f32ls): the pair loops of "
                    "kernel 1 on the FP32 pipe, everything else fp64; same workload as the headline line",
            "ms_per_step": ms32, "value": P * F * A / (ms32 * 1e-3), "unit": UNIT,
            "k_absorption_ms": k32_ms, "k_absorption_ms_fp64": k_ms,
            "max_abs_dtb_vs_fp64_path_K": dtb, "max_rel_dtau_vs_fp64_path": dtau, "gate_K": 0.01,
            "timing": "one CUDA-event pair around %d steps, L2 flush inside" % n32}
        del plan32, out32
        # the one-line import swap: the reference's own call pattern, one
        # TbCloudRTE object per (profile, elevation) with ONE angle, host arrays in, DataFrame out
        if rank == 0:
            prof = [d[k][0] for k in keys]

            def one_execute():
                rte = pkg.TbCloudRTE(prof[0], prof[1], prof[2], prof[3], frq, np.array([19.2]))
                rte.init_absmdl(args.model)
                rte.satellite = False
                return rte.execute()["tbtotal"].values
            for _ in range(5):
                one_execute()
            t0 = time.perf_counter()
            ncall = 200
            for _ in range(ncall):
                tb1 = one_execute()
            dt = (time.perf_counter() - t0) / ncall
            extra["class_api"] = {
                "what": "TbCloudRTE(z,p,t,rh,frq,[angle]); init_absmdl; satellite=False; execute() -- the "
                        "reference's per-call pattern (PyRTlib_processing.py:123-127) through the drop-in class, "
                        "wall clock per call incl. H2D, two kernels, D2H and the DataFrame",
                "ms_per_execute": dt * 1e3, "tb_per_s": len(frq) / dt, "calls": ncall,
                "finite": bool(np.isfinite(tb1).all())}
        cfgs = []
        for cfg in (3, 5, 4):
            try:
                cfgs.append(run_config(pkg, cfg, rank, world, reps=3, scale=args.scale,
                                       gather="nccl" if args.gather == "nccl" else "push"))
            except Exception as exc:      # e.g. out of memory on a smaller part: say so, keep the line
                cfgs.append({"config": CONFIG_NAMES[cfg], "error": repr(exc)})
                torch.cuda.empty_cache()
        extra["configs"] = cfgs

    if rank != 0:
        if N > 1:
            dist.destroy_process_group()
        return
    clocks = sampler.summary(t_wall0, t_wall1)
    # ---- cpu_baseline leg: the PyRTlib-shaped port timed on the host cores (N=1 only) and the oracle as
    # the CHECKER of this run's GPU results on 256 seeded profiles (outside every timed region)
    cpu, parity = None, None
    os.sched_setaffinity(0, orig_affinity)      # the CPU arm may use every host core again
    if not args.no_cpu_baseline:
        with tempfile.TemporaryDirectory() as tmp:
            pfile = os.path.join(tmp, "parity.npz")
            if N == 1:
                c = cpu_arm(1, 0, budget_s=15.0, parity_out=pfile)
                cpu = {"value": c["value"], "unit": UNIT, "cores": c["cores"], "kind": c["kind"],
                       "sample": c["sample"], "s_per_execute": c["s_per_execute"]}
            else:
                cpu_arm(1, 0, budget_s=0.5, parity_out=pfile)
            ref = np.load(pfile)
            dp = pkg.synth_profiles(int(ref["nprof"]), int(ref["levels"]), seed=int(ref["seed"]))
            got = pkg.simulate(dp["z"], dp["p"], dp["t"], dp["rh"], frq, ang, args.model)
            parity = {"profiles": int(ref["nprof"]),
                      "max_abs_dtb_K": {k: float(np.max(np.abs(got[k].cpu().numpy() - ref[k])))
                                        for k in ("tbtotal", "tbatm", "tmr")},
                      "max_rel_dtau": {k: float(np.max(np.abs(got[k].cpu().numpy() / ref[k] - 1.0)))
                                       for k in ("tauwet", "taudry")},
                      "oracle": "oracle/pyrtlib_vec.py (in-repo restatement; parity with real PyRTlib unpinned)"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": N, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "model": args.model, "profiles_per_gpu": P, "levels": L,
                       "channels": F, "elevations": A, "parallelism": "profiles sharded by rank (dp%d)" % N,
                       "outputs": list(fields),
                       "collective": gather_desc,
                       "l2": "256 MiB flush before every timed step, INSIDE the timed region "
                             "(%.3f ms per flush measured separately, not subtracted)" % flush_ms,
                       "timing": "one CUDA-event pair around all K steps incl. flushes and the tail of the last "
                                 "gather, max over ranks"},
            "clocks": clocks,
            "flush_ms_per_step": flush_ms,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "api": "Plan.forward_host -> mwrt_plan_forward_host (pinned host buffers)",
                    "numa": numa,
                    "fields_moved": moved,
                    "note": "all six outputs delivered to host arrays; tauliq (identically 0 in clear sky) "
                            "is filled by the host while the GPU works instead of crossing PCIe",
                    "tbtotal_only": {"value": e2e1_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                                     "d2h_bytes_per_step": 8 * P * F * A + 4 * P,
                                     "note": "same call asking for `tbtotal` alone, the one column the "
                                             "reference's driver reads (PyRTlib_processing.py:127)"}},
            "gpu_launches": args.steps * plan.launch_count(P),
            "roofline": roof, "cpu_baseline": cpu, "checksum_tbtotal": checksum,
            "max_abs_dtb_vs_oracle_K": None if parity is None else parity["max_abs_dtb_K"]["tbtotal"],
            "parity": parity, "extra": extra}
    _emit(line)
    if N > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
