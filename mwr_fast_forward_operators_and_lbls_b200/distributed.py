"""Multi-GPU plumbing: one process per GPU, profiles sharded by rank, ONE gather at the end.

The path is embarrassingly parallel over profiles -- every `TbCloudRTE` call of the reference is
independent (`/root/reference/python_src/proc/PyRTlib_processing.py:99-151` has no cross-iteration
state) -- so there is no collective inside the computation.  `torch.distributed`
(NCCL over NVLink 5 / NVSwitch on the GPU box, gloo in the CPU tests) is used only to assemble the
(P,F,A) results.
"""
import torch
import torch.distributed as dist


def shard_range(P, rank, world):
    """Contiguous block of profiles owned by `rank`: [start, stop).  Blocks differ by at most one."""
    base, rem = divmod(int(P), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_sizes(P, world):
    return [shard_range(P, r, world)[1] - shard_range(P, r, world)[0] for r in range(world)]


def gather_profiles_axis(local, P_total, group=None):
    """All-gather a tensor sharded along axis 0 by `shard_range` -> (P_total, ...) on every rank.
    One collective: shards are padded to the largest block and trimmed after."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    if world == 1:
        return local
    sizes = shard_sizes(P_total, world)
    pmax = max(sizes)
    if local.shape[0] != sizes[dist.get_rank(group)]:
        raise ValueError("local shard has %d profiles, expected %d" % (local.shape[0],
                                                                       sizes[dist.get_rank(group)]))
    tail = tuple(local.shape[1:])
    if local.shape[0] != pmax:
        pad = torch.zeros((pmax - local.shape[0],) + tail, dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    local = local.contiguous()
    full = torch.empty((world * pmax,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(full, local, group=group)
    if all(s == pmax for s in sizes):
        return full
    return torch.cat([full[r * pmax:r * pmax + sizes[r]] for r in range(world)], dim=0)


def gather_outputs(out, P_total, fields=None, group=None):
    """Gather a dict of per-rank (P_r,F,A) tensors.  The fields are stacked into one buffer so that a
    single collective moves them all."""
    fields = list(fields or [k for k in out if k != "status"])
    stacked = torch.stack([out[k] for k in fields], dim=1)          # (P_r, nfield, F, A)
    full = gather_profiles_axis(stacked, P_total, group)
    return {k: full[:, i] for i, k in enumerate(fields)}


def simulate_sharded(plan, z, p, t, rh, denliq=None, denice=None, fields=None, group=None):
    """Every rank passes the FULL (P,L) device tensors (or identical copies); each computes its block
    and all ranks return the assembled (P,F,A) results."""
    P = z.shape[0]
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    a, b = shard_range(P, rank, world)
    sl = slice(a, b)
    kw = {}
    if fields is not None:
        kw["fields"] = tuple(fields)
    out = plan.forward(z[sl].contiguous(), p[sl].contiguous(), t[sl].contiguous(), rh[sl].contiguous(),
                       None if denliq is None else denliq[sl].contiguous(),
                       None if denice is None else denice[sl].contiguous(), **kw)
    return gather_outputs(out, P, fields=fields or [k for k in out if k not in ("status", "tau_layer")],
                          group=group)
