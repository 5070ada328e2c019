#!/usr/bin/env python3
"""Multi-GPU check of the copy-engine gather (CopyEngineGather: one cudaMemcpyAsync per peer onto
symmetric memory, no collective kernel) against the NCCL all-gather (PipelinedForward): the gathered
fields must be bit-identical on every rank over several pipelined steps with DIFFERENT inputs per step
(a stale or torn buffer shows up as a mismatch), and the per-step time of both is printed.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/check_ce_gather.py
"""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist
import mwr_fast_forward_operators_and_lbls_b200 as pkg
from mwr_fast_forward_operators_and_lbls_b200.distributed import CopyEngineGather, PipelinedForward

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
P, L = int(os.environ.get("MWRT_CHECK_P", "10000")), 100
fields = ("tbtotal", "tbatm", "tmr", "tauwet", "taudry", "tauliq")
plan = pkg.Plan("R24", L, pkg.HATPRO_FRQ, pkg.HATPRO_ELEV)
batches = []
for s in range(3):       # three different batches, cycled
    d = pkg.synth_profiles(P, L, seed=100 * s + rank)
    if s == 1 and rank == world - 1:
        d["t"][17, 5] = float("nan")          # a flagged profile: NaN rows must arrive everywhere
    batches.append({k: torch.from_numpy(d[k]).cuda() for k in ("z", "p", "t", "rh")})
ref = PipelinedForward(plan, P, fields)
try:
    ce = CopyEngineGather(plan, P, fields)
except Exception as exc:      # no symmetric memory on this box
    if rank == 0:
        print("copy-engine gather unavailable: %r" % (exc,), flush=True)
    dist.destroy_process_group()
    sys.exit(77)
ok = True
slots = []
for s in range(7):            # pipelined: check the result of step s after step s+1 was queued
    b = batches[s % 3]
    sr = ref.run(b["z"], b["p"], b["t"], b["rh"])
    sc = ce.run(b["z"], b["p"], b["t"], b["rh"])
    slots.append((sr, sc))
    if s >= 1:
        # step s-1 landed?  wait for it explicitly (the events of that slot), then compare every field
        ref.gathered_ev[slots[s - 1][0]].synchronize()
        ce.landed[slots[s - 1][1]].synchronize()
        for f in fields:
            a = ref.field(f, slots[s - 1][0])
            c = ce.field(f, slots[s - 1][1])
            same = bool(torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(c, nan=-7.0)))
            ok = ok and same
ref.finish(); ce.finish()
torch.cuda.synchronize(); dist.barrier()
allsame = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(allsame, op=dist.ReduceOp.MIN)
if rank == 0:
    print("world %d: copy-engine gather identical to NCCL on every rank, every step: %s"
          % (world, bool(allsame.item())), flush=True)


def timeit(obj, steps=50):
    b = batches[0]
    for _ in range(5):
        obj.run(b["z"], b["p"], b["t"], b["rh"])
    obj.finish(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        obj.run(b["z"], b["p"], b["t"], b["rh"])
    obj.finish()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


if not os.environ.get("MWRT_CHECK_NOTIME"):
    t_ref, t_ce = timeit(ref), timeit(ce)
    # single-GPU-equivalent step for reference: the plan alone, no gather
    out = {}
    b = batches[0]
    for _ in range(5):
        out = plan.forward(b["z"], b["p"], b["t"], b["rh"], fields=fields, out=out)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        out = plan.forward(b["z"], b["p"], b["t"], b["rh"], fields=fields, out=out)
    e1.record(); torch.cuda.synchronize()
    t0 = e0.elapsed_time(e1) / 50
    if rank == 0:
        print("ms/step: no gather %.3f | NCCL all-gather (pipelined) %.3f | copy engines %.3f"
              % (t0, t_ref, t_ce), flush=True)
dist.destroy_process_group()
sys.exit(0 if allsame.item() else 1)
