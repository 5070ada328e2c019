#!/usr/bin/env python3
"""Why does the host entry point not scale to 8 ranks?  
Every rank, at the same time (barrier-released): pinned H2D / D2H bandwidth of the bench's buffer
sizes and the `forward_host` step -- once with the inherited CPU affinity and once bound to the GPU's
NUMA node (hostbind.bind_to_gpu_node BEFORE the pinned buffers are allocated, so first touch puts the
pages on that node).  Rank 0 prints one JSON line per setting with the per-rank figures.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29541 tools/e2e_diag.py            # run twice: MWRT_NO_BIND=1 and without
"""
import json
import os
import sys
import time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist
import mwr_fast_forward_operators_and_lbls_b200 as pkg
from mwr_fast_forward_operators_and_lbls_b200 import hostbind

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
info = {"bound": False, "numa_node": hostbind.gpu_numa_node(local)} if os.environ.get("MWRT_NO_BIND") \
    else hostbind.bind_to_gpu_node(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


P, L, F, A = 10000, 100, 14, 10
fields = ("tbtotal", "tbatm", "tmr", "tauwet", "taudry", "tauliq")
d = pkg.synth_profiles(P, L, seed=1234 + rank)
keys = ("z", "p", "t", "rh")
hin = pkg.pinned_fields(keys, (P, L))
for k in keys:
    hin[k][...] = d[k]
hout = pkg.pinned_fields(fields, (P, F, A))
hout["status"] = torch.empty((P,), dtype=torch.int32).pin_memory().numpy()
plan = pkg.Plan("R24", L, pkg.HATPRO_FRQ, pkg.HATPRO_ELEV)
# raw copies of the same sizes: 32 MB up, 56 MB down
up_h = torch.empty(4 * P * L, dtype=torch.float64).pin_memory()
up_d = torch.empty_like(up_h, device="cuda")
dn_h = torch.empty(5 * P * F * A, dtype=torch.float64).pin_memory()
dn_d = torch.empty_like(dn_h, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, n=20):
    for _ in range(3):
        fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


def both():
    with torch.cuda.stream(s1):
        up_d.copy_(up_h, non_blocking=True)
    with torch.cuda.stream(s2):
        dn_h.copy_(dn_d, non_blocking=True)
    s1.synchronize(); s2.synchronize()


res = {"h2d_gbs": up_h.numel() * 8 / timed(lambda: (up_d.copy_(up_h, non_blocking=True), torch.cuda.synchronize())) / 1e9,
       "d2h_gbs": dn_h.numel() * 8 / timed(lambda: (dn_h.copy_(dn_d, non_blocking=True), torch.cuda.synchronize())) / 1e9,
       "both_ms": timed(both) * 1e3,
       "forward_host_ms": timed(lambda: plan.forward_host(hin["z"], hin["p"], hin["t"], hin["rh"], fields=fields, out=hout)) * 1e3,
       "numa_node": info.get("numa_node"), "bound": info.get("bound"), "cpus": len(os.sched_getaffinity(0))}
allres = [None] * world
if world > 1:
    dist.all_gather_object(allres, res)
else:
    allres = [res]
if rank == 0:
    def col(k):
        return [round(r[k], 3) if isinstance(r[k], float) else r[k] for r in allres]
    print(json.dumps({"ranks": world, "bind": not bool(os.environ.get("MWRT_NO_BIND")),
                      "h2d_gbs": col("h2d_gbs"), "d2h_gbs": col("d2h_gbs"), "both_ms": col("both_ms"),
                      "forward_host_ms": col("forward_host_ms"), "numa_node": col("numa_node"),
                      "bound": col("bound"), "cpus": col("cpus"),
                      "sum_gbs_both": round(sum((up_h.numel() + dn_h.numel()) * 8 / (r["both_ms"] * 1e-3) / 1e9
                                                for r in allres), 1)}), flush=True)
if world > 1:
    dist.destroy_process_group()
