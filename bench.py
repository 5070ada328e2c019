#!/usr/bin/env python3
"""Headline benchmark: profile*channel*angle brightness temperatures per second (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

Workload (config.workload): BASELINE.json configs[1] -- 10^4 synthetic radiosonde-shaped profiles
(100 levels) x 14 HATPRO channels x 10 elevations, clear sky, model R24 (the reference's analysis
baseline, PyRTlib_processing.py:128-135), PER GPU (weak scaling: profiles shard by rank, the only
collective is the final all-gather of the results).  One "step" = one pass of the whole path
(vapour pressure -> absorption -> slant path -> RTE) over that batch.

  value  device-resident inputs, CUDA-event time summed over the K steps, max over ranks
  e2e    the same batch through the host entry point (mwrt_plan_forward_host): pinned host buffers in,
         host buffers out, H2D and D2H copies inside the timed region
  roofline  dominant kernel = k_absorption, timed alone with CUDA events; fp64-pipe bound, so the
         denominator is a DFMA-stream peak measured in the same run (MEASURED_PEAKS.json has no fp64
         figure); HBM numbers are given beside it
  cpu_baseline  oracle/cpu_baseline.py (PyRTlib-shaped port, all host cores) on a bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "profile*channel*angle TBs/sec"
UNIT = "TB/s (brightness temperatures per second)"
P_PER_GPU, LEVELS, MODEL = 10000, 100, "R24"
WORKLOAD = ("BASELINE.json configs[1]: 10^4 synthetic radiosonde-shaped profiles (100 levels) x 14 "
            "HATPRO channels x 10 elevations, clear sky, R24, per GPU")


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


def cpu_arm(steps, warmup, budget_s):
    cmd = [sys.executable, "-m", "oracle.cpu_baseline", "--levels", str(LEVELS), "--model", MODEL,
           "--steps", str(steps), "--warmup", str(warmup), "--budget-s", str(budget_s)]
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--model", default=MODEL)
    ap.add_argument("--gather", default="nccl", choices=["fused", "nccl"],
                    help="N>1: 'fused' = kernel 3 stores its results straight into every rank's gathered "
                         "buffer over NVLink (multicast / peer stores, no collective kernel); 'nccl' = "
                         "all_gather_into_tensor overlapped with the next step")
    ap.add_argument("--split", type=int, default=1,
                    help="mwrt_plan_set_split: pieces of the device-entry batch (kernel 3 of piece c on a side "
                         "stream while kernel 1 of piece c+1 runs); spreads the fused gather's stores in time")
    ap.add_argument("--gather-chunks", type=int, default=0,
                    help="N>1: 0 = gather of step s overlaps the kernels of step s+1 (default); k>0 = "
                         "within-step overlap over k pieces of the local shard")
    args = ap.parse_args()
    _reserve_stdout()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    import mwr_fast_forward_operators_and_lbls_b200 as pkg
    from mwr_fast_forward_operators_and_lbls_b200 import _lib, roofline
    from mwr_fast_forward_operators_and_lbls_b200.distributed import (FusedGatherForward, OverlappedForward,
                                                                     PipelinedForward)
    from mwr_fast_forward_operators_and_lbls_b200.forward import _ptr, _stream

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N = world
    P, L = P_PER_GPU, LEVELS
    frq, ang = pkg.HATPRO_FRQ, pkg.HATPRO_ELEV
    F, A = len(frq), len(ang)

    d = pkg.synth_profiles(P, L, seed=1234 + rank)
    keys = ("z", "p", "t", "rh")
    dev = {k: torch.from_numpy(d[k]).cuda() for k in keys}
    plan = pkg.Plan(args.model, L, frq, ang)
    if args.split > 1:
        plan.set_split(args.split)
    fields = ("tbtotal", "tbatm", "tmr", "tauwet", "taudry", "tauliq")   # north_star's six outputs
    out = {k: torch.empty((P, F, A), dtype=torch.float64, device="cuda") for k in fields}
    out["status"] = torch.empty((P,), dtype=torch.int32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # > 126 MB L2

    # N>1: the steps are a stream of batches; the all-gather of step s (six outputs, one collective)
    # runs on a side stream while step s+1 computes (double-buffered).  --gather-chunks > 0 selects the
    # within-step form instead (chunk c+1 computed while chunk c travels).
    ovl = None
    gather_desc = "none"
    if N > 1:
        if args.gather_chunks > 0:
            ovl = OverlappedForward(plan, P, fields, nchunk=args.gather_chunks)
            gather_desc = ("one all_gather_into_tensor of the six stacked outputs per chunk, %d chunks "
                           "overlapped within the step" % args.gather_chunks)
        else:
            if args.gather == "fused":
                try:
                    ovl = FusedGatherForward(plan, P, fields)
                    gather_desc = ("fused into kernel 3: every rank's RTE kernel stores its results into the "
                                   "gathered buffer of all ranks over NVLink (%s), one symmetric-memory "
                                   "barrier per step; no collective kernel" %
                                   ("multimem.st multicast through NVSwitch" if ovl.mode == "multicast"
                                    else "one store per peer"))
                except Exception as exc:      # symmetric memory not available: plain NCCL
                    if rank == 0:
                        print("fused gather unavailable (%s); using NCCL" % exc, file=sys.stderr)
                    ovl = None
            if ovl is None:
                ovl = PipelinedForward(plan, P, fields)
                gather_desc = ("one all_gather_into_tensor of the six stacked outputs per step, overlapped "
                               "with the next step's kernels (double-buffered); the last step's gather is "
                               "inside the timed total")

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
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    t_wall0 = time.time()
    for e0, e1 in ev:
        flush.zero_()          # L2 flush between timed steps, outside the event pair
        e0.record()
        res = step()
        e1.record()
    # tail: the last step's gather must have landed before the clock stops
    tail0, tail1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tail0.record()
    if isinstance(ovl, (PipelinedForward, FusedGatherForward)):
        ovl.finish()
    tail1.record()
    barrier()
    t_wall1 = time.time()
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev] + [tail0.elapsed_time(tail1)]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device="cuda")
    if N > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = N * P * F * A * args.steps / (total_ms * 1e-3)
    if isinstance(ovl, (PipelinedForward, FusedGatherForward)):
        checksum = float(ovl.field("tbtotal", res).sum().item())
    elif N > 1:
        checksum = float(res.field("tbtotal").sum().item())
    else:
        checksum = float(res["tbtotal"].sum().item())

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
    t_wall2 = time.time()

    # ---- dominant kernel alone (k_absorption) for the roofline
    lib = _lib.load()
    m = pkg.get_model(args.model)
    e_t = pkg.vapor(dev["t"], dev["rh"])[0]
    f_t = torch.from_numpy(frq).cuda()
    awet = torch.empty((P, F, L), dtype=torch.float64, device="cuda")
    adry = torch.empty_like(awet)

    def k_abs():
        _lib.check(lib.mwrt_absorption_f64(m.handle, P, L, F, _ptr(dev["p"]), _ptr(dev["t"]), _ptr(e_t),
                                           None, None, _ptr(f_t), _ptr(awet), _ptr(adry), None, None,
                                           None, _stream()), "absorption")
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
    k_ms = sum(kt) / len(kt)
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
    traffic = None   # measured DRAM bytes per launch from the committed ncu --set full capture
    try:
        if args.model == MODEL:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))["k_absorption"]["traffic_bytes"]
    except Exception:
        pass
    roof = {"bound": "fp64", "kernel": "k_absorption", "achieved": achieved,
            "peak": peak_flops / 1e12, "unit": "TFLOP/s", "frac": achieved / (peak_flops / 1e12),
            "peak_source": "self-measured DFMA stream in this run (nominal 37.2); MEASURED_PEAKS.json "
                           "has no fp64 figure",
            "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram read+write, profiles/r01_traffic.json)",
            "kernel_ms": k_ms,
            "kernel_share_of_step": k_ms / (total_ms / args.steps),
            "algorithmic_flops_per_launch": w["flops_k_abs"],
            "hbm": {"algorithmic_bytes_per_launch": w["bytes_k_abs"],
                    "achieved_gbs": w["bytes_k_abs"] / (k_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
            "whole_step_tflops": w["flops_total"] / (total_ms / args.steps * 1e-3) / 1e12}

    if rank != 0:
        if N > 1:
            dist.destroy_process_group()
        return
    clocks = sampler.summary(t_wall0, t_wall1)
    cpu = None
    if N == 1 and not args.no_cpu_baseline:
        c = cpu_arm(1, 0, budget_s=15.0)
        cpu = {"value": c["value"], "unit": UNIT, "cores": c["cores"], "kind": c["kind"],
               "sample": c["sample"], "s_per_execute": c["s_per_execute"]}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": N, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "model": args.model, "profiles_per_gpu": P, "levels": L,
                       "channels": F, "elevations": A, "parallelism": "profiles sharded by rank (dp%d)" % N,
                       "outputs": list(fields),
                       "collective": gather_desc,
                       "l2": "256 MiB flush between timed steps, outside the CUDA-event pairs",
                       "timing": "sum of per-step CUDA-event intervals, max over ranks"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "api": "Plan.forward_host -> mwrt_plan_forward_host (pinned host buffers)",
                    "fields_moved": moved,
                    "note": "all six outputs delivered to host arrays; tauliq (identically 0 in clear sky) "
                            "is filled by the host while the GPU works instead of crossing PCIe",
                    "tbtotal_only": {"value": e2e1_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                                     "d2h_bytes_per_step": 8 * P * F * A + 4 * P,
                                     "note": "same call asking for `tbtotal` alone, the one column the "
                                             "reference's driver reads (PyRTlib_processing.py:127)"}},
            "gpu_launches": args.steps * (plan.launch_count(P) * (args.gather_chunks if (N > 1 and args.gather_chunks > 0) else 1)),
            "roofline": roof, "cpu_baseline": cpu, "checksum_tbtotal": checksum,
            "max_abs_dtb_vs_oracle_K": "see tests/test_gpu_parity.py (1e-3 K gate) and profiles/r01_parity.jsonl (measured <= 2e-10 K; absorption <= 1.4e-12 relative)"}
    _emit(line)
    if N > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
