#!/usr/bin/env python3
"""Multi-GPU check of the gather fused into kernel 3 (FusedGatherForward) against the NCCL
all-gather (PipelinedForward): identical gathered results on every rank, and the per-step time of both.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/check_fused_gather.py [mode]      # mode: auto | multicast | peers
"""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import torch.distributed as dist
import mwr_fast_forward_operators_and_lbls_b200 as pkg
from mwr_fast_forward_operators_and_lbls_b200.distributed import FusedGatherForward, PipelinedForward

mode = sys.argv[1] if len(sys.argv) > 1 else "auto"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
P, L = int(os.environ.get("MWRT_CHECK_P", "10000")), 100
d = pkg.synth_profiles(P, L, seed=1234 + rank)
dev = {k: torch.from_numpy(d[k]).cuda() for k in ("z", "p", "t", "rh")}
fields = ("tbtotal", "tbatm", "tmr", "tauwet", "taudry", "tauliq")
plan = pkg.Plan("R24", L, pkg.HATPRO_FRQ, pkg.HATPRO_ELEV)
ref = PipelinedForward(plan, P, fields)
try:
    fus = FusedGatherForward(plan, P, fields, mode=mode)
except Exception as exc:      # no symmetric memory / multicast on this box
    if rank == 0:
        print("fused gather unavailable: %r" % (exc,), flush=True)
    dist.destroy_process_group()
    sys.exit(77)
if rank == 0:
    print("world %d, fused mode: %s" % (world, fus.mode), flush=True)
s0 = ref.run(dev["z"], dev["p"], dev["t"], dev["rh"]); ref.finish()
b0 = fus.run(dev["z"], dev["p"], dev["t"], dev["rh"])
torch.cuda.synchronize(); dist.barrier()
same = bool(torch.equal(ref.gathered[s0], fus.gathered[b0]))
allsame = torch.tensor([1 if same else 0], device="cuda")
dist.all_reduce(allsame, op=dist.ReduceOp.MIN)
if rank == 0:
    print("gathered buffers identical on every rank:", bool(allsame.item()), flush=True)


def timeit(obj, steps=50):
    for _ in range(5):
        obj.run(dev["z"], dev["p"], dev["t"], dev["rh"])
    obj.finish(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        obj.run(dev["z"], dev["p"], dev["t"], dev["rh"])
    obj.finish()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


if os.environ.get("MWRT_CHECK_NOTIME"):
    dist.destroy_process_group()
    sys.exit(0 if allsame.item() else 1)
t_ref, t_fus = timeit(ref), timeit(fus)
if rank == 0:
    print("ms/step: NCCL all-gather (pipelined) %.3f | fused into kernel 3 (%s) %.3f" % (t_ref, fus.mode, t_fus), flush=True)
dist.destroy_process_group()
sys.exit(0 if allsame.item() else 1)
