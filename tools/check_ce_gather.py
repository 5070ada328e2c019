#!/usr/bin/env python3
"""Multi-GPU check of the push gather (PushGather: each rank's slot of a symmetric buffer delivered by
a small NVLink-multicast kernel, by per-peer stores, or by one cudaMemcpyAsync per peer; no collective
call) against the NCCL all-gather (PipelinedForward): the gathered
fields must be bit-identical on every rank over several pipelined steps with DIFFERENT inputs per step
(a stale or torn buffer shows up as a mismatch), and the per-step time of both is printed.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/check_ce_gather.py [auto|multicast|peers|ce]
"""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist
import mwr_fast_forward_operators_and_lbls_b200 as pkg
from mwr_fast_forward_operators_and_lbls_b200.distributed import PipelinedForward, PushGather

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
transport = sys.argv[1] if len(sys.argv) > 1 else "auto"      # auto | multicast | peers | ce
ref = PipelinedForward(plan, P, fields)
try:
    ce = PushGather(plan, P, fields, transport=transport)
except Exception as exc:      # no symmetric memory / multicast on this box
    if rank == 0:
        print("push gather (%s) unavailable: %r" % (transport, exc), flush=True)
    dist.destroy_process_group()
    sys.exit(77)
if rank == 0:
    print("push gather transport: %s, %d CTAs" % (ce.transport, ce.nblocks), flush=True)
# Pipelined use under the consumers' contract (PushGather docstring): run(s), then consume step s-1 ON THE
# STREAM (here: a snapshot copy), so that the consumer of buffer b is queued before the run() that reuses
# b.  Ranks are skewed on purpose (host sleeps that differ by rank and step): a peer that runs ahead must
# not be able to overwrite a buffer this rank has not consumed yet.
import random
import time
random.seed(1000 + rank)
nstep = int(os.environ.get("MWRT_CHECK_STEPS", "12"))
snaps_ref, snaps_ce, slots = [], [], []


def consume(step):
    sr, sc = slots[step]
    torch.cuda.current_stream().wait_event(ref.gathered_ev[sr])
    ce.wait(sc)
    snaps_ref.append(torch.stack([ref.field(f, sr) for f in fields]).clone())
    snaps_ce.append(torch.stack([ce.field(f, sc) for f in fields]).clone())


for s in range(nstep):
    b = batches[s % 3]
    time.sleep(random.uniform(0.0, 0.02) if (s + rank) % 3 else 0.0)
    slots.append((ref.run(b["z"], b["p"], b["t"], b["rh"]), ce.run(b["z"], b["p"], b["t"], b["rh"])))
    if s >= 1:
        consume(s - 1)
consume(nstep - 1)
ref.finish(); ce.finish()
torch.cuda.synchronize(); dist.barrier()
ok = all(bool(torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(c, nan=-7.0)))
         for a, c in zip(snaps_ref, snaps_ce))
nanrow = bool(torch.isnan(snaps_ce[1][0, (world - 1) * P + 17]).all().item())   # the flagged profile of batch 1
ok = ok and nanrow
del snaps_ref, snaps_ce
allsame = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(allsame, op=dist.ReduceOp.MIN)
if rank == 0:
    print("world %d: push gather (%s) identical to NCCL on every rank, every step: %s"
          % (world, ce.transport, bool(allsame.item())), flush=True)


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
        print("ms/step: no gather %.3f | NCCL all-gather (pipelined) %.3f | push gather (%s) %.3f"
              % (t0, t_ref, ce.transport, t_ce), flush=True)
dist.destroy_process_group()
sys.exit(0 if allsame.item() else 1)
