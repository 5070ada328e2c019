"""Multi-GPU plumbing: one process per GPU, profiles sharded by rank, ONE gather at the end.

The path is embarrassingly parallel over profiles -- every `TbCloudRTE` call of the reference is
independent (`/root/reference/python_src/proc/PyRTlib_processing.py:99-151` has no cross-iteration
state) -- so there is no collective inside the computation.  `torch.distributed`
(NCCL over NVLink 5 / NVSwitch on the GPU box, gloo in the CPU tests) is used only to assemble the
(P,F,A) results.
"""
import ctypes as C

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


class OverlappedForward:
    """Compute the local shard in `nchunk` pieces and all-gather each piece on a side stream while
    the next piece is being computed, so the final gather hides behind the kernels.

    Result layout: `gathered[c]` is a (world, Pc, nfield, F, A) tensor holding chunk c of every
    rank; `field(name)` assembles the contiguous (P_total, F, A) view order (rank-major, then local
    index) on demand.  Equal shard sizes are required (weak-scaling batches); use `simulate_sharded`
    for ragged shards.
    """

    def __init__(self, plan, P_local, fields, nchunk=2, group=None):
        self.plan, self.fields, self.group = plan, tuple(fields), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.P_local = int(P_local)
        self.nchunk = max(1, min(int(nchunk), self.P_local))
        base, rem = divmod(self.P_local, self.nchunk)
        self.bounds = []
        a = 0
        for c in range(self.nchunk):
            b = a + base + (1 if c < rem else 0)
            self.bounds.append((a, b))
            a = b
        nf, F, A = len(self.fields), plan.Fo, plan.A
        dev = torch.device("cuda", torch.cuda.current_device())
        # local results, field-stacked per chunk so that one collective moves all fields of a chunk
        self.local = [torch.empty((nf, b - a, F, A), dtype=torch.float64, device=dev)
                      for a, b in self.bounds]
        self.status = torch.empty((self.P_local,), dtype=torch.int32, device=dev)
        self.gathered = [torch.empty((self.world,) + tuple(t.shape), dtype=torch.float64, device=dev)
                         for t in self.local]
        self.comm = torch.cuda.Stream()
        self.events = [torch.cuda.Event() for _ in self.bounds]

    def run(self, z, p, t, rh, denliq=None, denice=None):
        main = torch.cuda.current_stream()
        self.comm.wait_stream(main)          # previous consumers of `gathered` are ordered before us
        for c, (a, b) in enumerate(self.bounds):
            out = {k: self.local[c][i] for i, k in enumerate(self.fields)}
            out["status"] = self.status[a:b]
            self.plan.forward(z[a:b], p[a:b], t[a:b], rh[a:b],
                              None if denliq is None else denliq[a:b],
                              None if denice is None else denice[a:b], fields=self.fields, out=out)
            self.events[c].record(main)
            if self.world > 1:
                with torch.cuda.stream(self.comm):
                    self.comm.wait_event(self.events[c])
                    dist.all_gather_into_tensor(self.gathered[c], self.local[c], group=self.group)
            else:
                self.gathered[c] = self.local[c].unsqueeze(0)
        main.wait_stream(self.comm)
        return self

    def field(self, name):
        """(P_total, F, A) tensor of one field, profiles ordered rank-major."""
        i = self.fields.index(name)
        return torch.cat([g[:, i] for g in self.gathered], dim=1).reshape(
            self.world * self.P_local, self.plan.Fo, self.plan.A)


class PipelinedForward:
    """Streaming form for a sequence of batches (steps): the all-gather of step s runs on a side
    stream WHILE step s+1 is being computed (double-buffered results), so in steady state a step
    costs max(compute, gather) instead of their sum.  `finish()` makes the current stream wait for
    the last gather; results of step s are in `gathered[s % 2]`, laid out (world, nfield, P_local, F, A).
    """

    def __init__(self, plan, P_local, fields, group=None):
        self.plan, self.fields, self.group = plan, tuple(fields), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.P_local = int(P_local)
        nf, F, A = len(self.fields), plan.Fo, plan.A
        dev = torch.device("cuda", torch.cuda.current_device())
        self.local = [torch.empty((nf, self.P_local, F, A), dtype=torch.float64, device=dev) for _ in range(2)]
        self.status = [torch.empty((self.P_local,), dtype=torch.int32, device=dev) for _ in range(2)]
        self.gathered = [torch.empty((self.world, nf, self.P_local, F, A), dtype=torch.float64, device=dev)
                         for _ in range(2)]
        self.comm = torch.cuda.Stream()
        self.computed = [torch.cuda.Event() for _ in range(2)]
        self.gathered_ev = [torch.cuda.Event() for _ in range(2)]
        self.step = 0

    def run(self, z, p, t, rh, denliq=None, denice=None):
        s = self.step % 2
        main = torch.cuda.current_stream()
        if self.step >= 2:
            main.wait_event(self.gathered_ev[s])      # slot s is free once its previous gather is done
        out = {k: self.local[s][i] for i, k in enumerate(self.fields)}
        out["status"] = self.status[s]
        self.plan.forward(z, p, t, rh, denliq, denice, fields=self.fields, out=out)
        self.computed[s].record(main)
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(self.computed[s])
            if self.world > 1:
                dist.all_gather_into_tensor(self.gathered[s], self.local[s], group=self.group)
            else:
                self.gathered[s][0].copy_(self.local[s])
            self.gathered_ev[s].record(self.comm)
        self.step += 1
        return s

    def finish(self):
        torch.cuda.current_stream().wait_stream(self.comm)

    def field(self, name, slot):
        i = self.fields.index(name)
        return self.gathered[slot][:, i].reshape(self.world * self.P_local, self.plan.Fo, self.plan.A)


class CopyEngineGather:
    """The gather on the COPY ENGINES: every rank pushes its finished fields straight into the gathered
    buffer of each peer with ONE asynchronous device-to-device copy per peer (`mwrt_push_async` =
    cudaMemcpyAsync onto peer-mapped symmetric memory, one side stream per peer), while the next batch
    is being computed.  No collective kernel runs, so nothing takes SMs from the two FP64-bound kernels
    (an NCCL all-gather occupies 16+ CTAs for the whole transfer); a symmetric-memory barrier per step,
    on the side stream, tells every rank that all pushes of the step have landed.

    Only what carries information travels: in clear sky `tauliq` / `tauice` are identically zero (NaN
    for flagged profiles) and are rebuilt from the gathered status words by `field()`.

    Kernel 3 writes its results directly into this rank's slot of the symmetric buffer, so there is no
    local staging copy.  Results of step s: `gathered[s % 2]`, laid out (world, nfield, P_local, F, A)
    on every rank (double-buffered); `status_g[s % 2]` is (world, P_local).

    Ordering: the pushes of step s+2 overwrite the peers' copy of buffer s % 2, so they wait for the
    barrier of step s+1 -- every rank arrives there only after the consumers of its step-s results
    were queued ahead of its step-s+1 kernels.
    """

    def __init__(self, plan, P_local, fields, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self._lib = _lib
        self.lib = _lib.load()
        self.plan, self.fields = plan, tuple(fields)
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.P_local = int(P_local)
        F, A = plan.Fo, plan.A
        zero = () if plan.cloudy else ("tauliq", "tauice")
        self.moved = tuple(f for f in self.fields if f not in zero)
        self.zero_fields = tuple(f for f in self.fields if f in zero)
        nf = len(self.moved)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        shape = (self.world, nf, self.P_local, F, A)
        self.gathered = [symm_mem.empty(shape, dtype=torch.float64, device=dev) for _ in range(2)]
        self.hdl = [symm_mem.rendezvous(g, self.group) for g in self.gathered]
        self.status_g = [symm_mem.empty((self.world, self.P_local), dtype=torch.int32, device=dev) for _ in range(2)]
        self.hdl_st = [symm_mem.rendezvous(g, self.group) for g in self.status_g]
        self.rank_bytes = nf * self.P_local * F * A * 8
        self.st_bytes = self.P_local * 4
        # addresses of THIS rank's slot inside every peer's buffers, as mapped into this process
        self.peers = [q for q in range(self.world) if q != self.rank]
        self.dst = [{q: int(self.hdl[b].buffer_ptrs[q]) + self.rank * self.rank_bytes for q in self.peers}
                    for b in range(2)]
        self.dst_st = [{q: int(self.hdl_st[b].buffer_ptrs[q]) + self.rank * self.st_bytes for q in self.peers}
                       for b in range(2)]
        self.push = {q: torch.cuda.Stream() for q in self.peers}
        self.comm = torch.cuda.Stream()
        self.computed = [torch.cuda.Event() for _ in range(2)]
        self.pushed = [{q: torch.cuda.Event() for q in self.peers} for _ in range(2)]
        self.landed = [torch.cuda.Event() for _ in range(2)]     # barrier of the step using buffer b passed
        self.step = 0

    def run(self, z, p, t, rh, denliq=None, denice=None):
        b = self.step % 2
        main = torch.cuda.current_stream()
        if self.step >= 2:
            main.wait_event(self.landed[b])       # my pushes out of slot b (two steps ago) are complete
        own = self.gathered[b][self.rank]         # (nf, P_local, F, A): kernel 3 writes here directly
        out = {k: own[i] for i, k in enumerate(self.moved)}
        out["status"] = self.status_g[b][self.rank]
        self.plan.forward(z, p, t, rh, denliq, denice, fields=self.moved, out=out)
        self.computed[b].record(main)
        src, src_st = own.data_ptr(), self.status_g[b][self.rank].data_ptr()
        for q in self.peers:
            st = self.push[q]
            st.wait_event(self.computed[b])
            if self.step >= 1:
                st.wait_event(self.landed[1 - b])  # peers are done with what buffer b held two steps ago
            h = C.c_void_p(st.cuda_stream)
            self._lib.check(self.lib.mwrt_push_async(C.c_void_p(self.dst[b][q]), C.c_void_p(src),
                                                     self.rank_bytes, h), "push_async")
            self._lib.check(self.lib.mwrt_push_async(C.c_void_p(self.dst_st[b][q]), C.c_void_p(src_st),
                                                     self.st_bytes, h), "push_async")
            self.pushed[b][q].record(st)
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(self.computed[b])
            for q in self.peers:
                self.comm.wait_event(self.pushed[b][q])
            self.hdl[b].barrier(channel=b)         # every rank's pushes of this step have landed everywhere
            self.landed[b].record(self.comm)
        self.step += 1
        return b

    def finish(self):
        torch.cuda.current_stream().wait_stream(self.comm)

    def field(self, name, slot):
        """(world * P_local, F, A) view of one gathered field, profiles ordered rank-major."""
        n = self.world * self.P_local
        if name in self.zero_fields:              # identically zero in clear sky; NaN where flagged
            out = torch.zeros((n, self.plan.Fo, self.plan.A), dtype=torch.float64, device=self.dev)
            bad = (self.status_g[slot].reshape(n) & 11) != 0      # NAN_INPUT | NEG_ABS | RANGE
            out[bad] = float("nan")
            return out
        i = self.moved.index(name)
        return self.gathered[slot][:, i].reshape(n, self.plan.Fo, self.plan.A)

    def status(self, slot):
        return self.status_g[slot].reshape(self.world * self.P_local)


class FusedGatherForward:
    """The gather fused into kernel 3: every rank's RTE kernel stores its results straight into the
    gathered buffer of ALL ranks -- with NVLink multicast stores (`multimem.st`: one store, NVSwitch
    replicates it into the same offset of every rank's symmetric buffer) or, where multicast is not
    available, with one plain store per peer through peer-mapped memory.  No collective kernel runs:
    nothing competes with the next step's kernels for SMs, and the 67 MB a rank produces leave over
    NVLink while kernel 3 is still integrating ("RTE epilogue writes through peer
    pointers").  A symmetric-memory barrier after the step orders the ranks.

    Results of step s: `gathered[s % 2]`, laid out (world, nfield, P_local, F, A) on every rank
    (double-buffered: step s+1 writes the other buffer).
    """

    def __init__(self, plan, P_local, fields, group=None, mode="auto"):
        import torch.distributed._symmetric_memory as symm_mem
        self.plan, self.fields = plan, tuple(fields)
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.P_local = int(P_local)
        nf, F, A = len(self.fields), plan.Fo, plan.A
        dev = torch.device("cuda", torch.cuda.current_device())
        shape = (self.world, nf, self.P_local, F, A)
        self.gathered = [symm_mem.empty(shape, dtype=torch.float64, device=dev) for _ in range(2)]
        self.hdl = [symm_mem.rendezvous(g, self.group) for g in self.gathered]
        self.status = [torch.empty((self.P_local,), dtype=torch.int32, device=dev) for _ in range(2)]
        mc = [int(getattr(h, "multicast_ptr", 0) or 0) for h in self.hdl]
        if mode == "auto":
            mode = "multicast" if all(mc) else "peers"
        if mode == "multicast" and not all(mc):
            raise RuntimeError("NVLink multicast is not available for this group")
        self.mode = mode
        self.field_bytes = self.P_local * F * A * 8
        self.rank_bytes = nf * self.field_bytes
        self.base = []          # per buffer: address kernel 3 is given for field 0 of this rank
        self.offsets = []       # per buffer: byte offsets of the peers' copies (mode "peers")
        for b in range(2):
            own = self.gathered[b].data_ptr()
            if mode == "multicast":
                self.base.append(mc[b] + self.rank * self.rank_bytes)
                self.offsets.append(None)
            else:
                ptrs = [int(q) for q in self.hdl[b].buffer_ptrs]
                self.base.append(own + self.rank * self.rank_bytes)
                self.offsets.append([q - own for q in ptrs])
        self.step = 0

    def run(self, z, p, t, rh, denliq=None, denice=None):
        b = self.step % 2
        # buffer b was last written two steps ago; every rank passed the barrier of the step in between
        # after its consumers of buffer b were queued, so it is free to be overwritten
        self.plan.set_output_fanout(self.mode, self.offsets[b])
        out = {k: self.base[b] + i * self.field_bytes for i, k in enumerate(self.fields)}
        out["status"] = self.status[b]
        try:
            self.plan.forward(z, p, t, rh, denliq, denice, fields=self.fields, out=out)
        finally:
            self.plan.set_output_fanout("plain")
        self.hdl[b].barrier(channel=b)     # all ranks' stores have landed everywhere (stream-ordered)
        self.step += 1
        return b

    def finish(self):
        pass

    def field(self, name, slot):
        i = self.fields.index(name)
        return self.gathered[slot][:, i].reshape(self.world * self.P_local, self.plan.Fo, self.plan.A)


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
