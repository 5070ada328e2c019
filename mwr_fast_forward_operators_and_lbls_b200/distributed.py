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


class PushGather:
    """The gather as a PUSH of each rank's finished fields into the gathered buffer of every peer,
    overlapped with the next batch's kernels -- no collective call, nothing but a barrier to wait for.

    Kernel 3 writes its results directly into this rank's slot of a symmetric buffer
    (`torch.distributed._symmetric_memory`); the slot then travels by one of three transports:

      "multicast"  `mwrt_push_kernel`: a handful of CTAs store the slot to the buffer's NVLink MULTICAST
                   address (`multimem.st`), NVSwitch replicates every store into all ranks' copies -- a
                   rank sends its 56 MB ONCE per step.  Default beyond 4 GPUs.
      "peers"      the same kernel with one plain store per peer-mapped copy (no multicast support).
      "ce"         one cudaMemcpyAsync per peer (`mwrt_push_async`) on per-peer side streams: the copy
                   engines.  Free at 2-4 GPUs, but seven concurrent peer copies reach only ~210 GB/s per
                   GPU on 8 x B200 (profiles/r02e_check_ce_gather_n8.log: 1.87 ms per step against 1.34
                   for NCCL and 1.14 of kernels).  Default up to 4 GPUs.

    Only what carries information travels: in clear sky `tauliq` / `tauice` are identically zero (NaN
    for flagged profiles) and are rebuilt from the gathered status words by `field()`.

    Results of step s: `gathered[s % 2]`, laid out (world, nfield, P_local, F, A) on every rank
    (double-buffered); `status_g[s % 2]` is (world, P_pad).  A symmetric-memory barrier per step, on
    the side stream, tells every rank that all pushes of the step have landed.

    Contract for consumers: the results of step s stay valid in `gathered[s % 2]` until `run()` is
    called for step s+2 -- consume (or copy) them on a CUDA stream that the CALLING stream of that
    `run()` is ordered after, typically the calling stream itself.  Mechanism: peers write into this
    rank's memory without its participation, so every `run()` starts with an ENTRY barrier (queued on
    the side stream behind an event recorded on the calling stream); the pushes into buffer b wait
    for it, i.e. until every rank's stream has passed the consumers it queued for what buffer b held.
    In steady state the entry barrier completes long before the step's kernels do and costs nothing.
    """

    def __init__(self, plan, P_local, fields, group=None, transport="auto", nblocks=0):
        import os
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self._lib = _lib
        self.lib = _lib.load()
        self.plan, self.fields = plan, tuple(fields)
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.P_local = int(P_local)
        self.P_pad = (self.P_local + 3) & ~3          # status rows padded to 16 bytes
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
        self.status_g = [symm_mem.empty((self.world, self.P_pad), dtype=torch.int32, device=dev) for _ in range(2)]
        self.hdl_st = [symm_mem.rendezvous(g, self.group) for g in self.status_g]
        for g in self.status_g:
            g.zero_()
        self.rank_bytes = nf * self.P_local * F * A * 8
        self.st_bytes = self.P_pad * 4
        mc = [int(getattr(h, "multicast_ptr", 0) or 0) for h in self.hdl + self.hdl_st]
        if transport == "auto":
            # measured on B200: up to 4 GPUs the copy engines deliver the slots for
            # free (no SM shared with the kernels); at 8 GPUs seven concurrent peer copies are too slow
            # and the multicast kernel, which sends every byte once, wins by 0.5 ms per step
            transport = "multicast" if (self.world > 4 and all(mc) and self.rank_bytes % 16 == 0) else "ce"
        if transport == "multicast" and not all(mc):
            raise RuntimeError("NVLink multicast is not available for this group")
        if transport in ("multicast", "peers") and self.rank_bytes % 16:
            raise RuntimeError("the push kernel moves 16-byte words; a rank's slot has %d bytes" % self.rank_bytes)
        self.transport = transport
        self.nblocks = int(nblocks or os.environ.get("MWRT_PUSH_BLOCKS", 0) or 8)
        # addresses of THIS rank's slot: inside every peer's buffers as mapped into this process, and
        # inside the multicast window
        self.peers = [q for q in range(self.world) if q != self.rank]
        self.dst = [{q: int(self.hdl[b].buffer_ptrs[q]) + self.rank * self.rank_bytes for q in self.peers}
                    for b in range(2)]
        self.dst_st = [{q: int(self.hdl_st[b].buffer_ptrs[q]) + self.rank * self.st_bytes for q in self.peers}
                       for b in range(2)]
        self.mc = [mc[b] + self.rank * self.rank_bytes for b in range(2)]
        self.mc_st = [mc[2 + b] + self.rank * self.st_bytes for b in range(2)]
        if transport == "ce":
            self.push = {q: torch.cuda.Stream() for q in self.peers}
        else:
            self.push = {None: torch.cuda.Stream(priority=-1)}     # one high-priority stream for the kernel
        self.comm = torch.cuda.Stream(priority=-1)
        self.computed = [torch.cuda.Event() for _ in range(2)]
        self.pushed = [{q: torch.cuda.Event() for q in self.push} for _ in range(2)]
        self.landed = [torch.cuda.Event() for _ in range(2)]     # barrier of the step using buffer b passed
        self.entered = [torch.cuda.Event() for _ in range(2)]    # calling stream reached run() for buffer b
        self.free = [torch.cuda.Event() for _ in range(2)]       # ... on EVERY rank: buffer b may be overwritten
        self.step = 0

    def _push_slot(self, b, src, src_st):
        chk, lib = self._lib.check, self.lib
        if self.transport == "ce":
            for q in self.peers:
                st = self.push[q]
                h = C.c_void_p(st.cuda_stream)
                chk(lib.mwrt_push_async(C.c_void_p(self.dst[b][q]), C.c_void_p(src), self.rank_bytes, h), "push_async")
                chk(lib.mwrt_push_async(C.c_void_p(self.dst_st[b][q]), C.c_void_p(src_st), self.st_bytes, h),
                    "push_async")
            return
        h = C.c_void_p(self.push[None].cuda_stream)
        if self.transport == "multicast":
            chk(lib.mwrt_push_kernel(C.c_void_p(src), self.rank_bytes, C.c_void_p(self.mc[b]), None, 0,
                                     self.nblocks, h), "push_kernel")
            chk(lib.mwrt_push_kernel(C.c_void_p(src_st), self.st_bytes, C.c_void_p(self.mc_st[b]), None, 0, 1, h),
                "push_kernel")
        else:
            n = len(self.peers)
            arr = (C.c_int64 * n)(*[self.dst[b][q] for q in self.peers])
            chk(lib.mwrt_push_kernel(C.c_void_p(src), self.rank_bytes, None, arr, n, self.nblocks, h), "push_kernel")
            arr = (C.c_int64 * n)(*[self.dst_st[b][q] for q in self.peers])
            chk(lib.mwrt_push_kernel(C.c_void_p(src_st), self.st_bytes, None, arr, n, 1, h), "push_kernel")

    def run(self, z, p, t, rh, denliq=None, denice=None):
        b = self.step % 2
        main = torch.cuda.current_stream()
        # entry barrier: every rank's calling stream is past the consumers of what buffer b held
        self.entered[b].record(main)
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(self.entered[b])
            self.hdl_st[b].barrier(channel=b)
            self.free[b].record(self.comm)
        if self.step >= 2:
            main.wait_event(self.landed[b])       # my pushes out of slot b (two steps ago) are complete
        own = self.gathered[b][self.rank]         # (nf, P_local, F, A): kernel 3 writes here directly
        out = {k: own[i] for i, k in enumerate(self.moved)}
        out["status"] = self.status_g[b][self.rank][:self.P_local]
        self.plan.forward(z, p, t, rh, denliq, denice, fields=self.moved, out=out)
        self.computed[b].record(main)
        for st in self.push.values():
            st.wait_event(self.computed[b])
            st.wait_event(self.free[b])            # no rank still reads what buffer b held
        self._push_slot(b, own.data_ptr(), self.status_g[b][self.rank].data_ptr())
        for q, st in self.push.items():
            self.pushed[b][q].record(st)
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(self.computed[b])
            for q in self.push:
                self.comm.wait_event(self.pushed[b][q])
            self.hdl[b].barrier(channel=b)         # every rank's pushes of this step have landed everywhere
            self.landed[b].record(self.comm)
        self.step += 1
        return b

    def wait(self, slot):
        """Order the current stream after the arrival of every rank's results in `gathered[slot]`."""
        torch.cuda.current_stream().wait_event(self.landed[slot])

    def finish(self):
        torch.cuda.current_stream().wait_stream(self.comm)

    def field(self, name, slot):
        """(world * P_local, F, A) view of one gathered field, profiles ordered rank-major."""
        n = self.world * self.P_local
        if name in self.zero_fields:              # identically zero in clear sky; NaN where flagged
            out = torch.zeros((n, self.plan.Fo, self.plan.A), dtype=torch.float64, device=self.dev)
            bad = (self.status(slot) & 11) != 0   # NAN_INPUT | NEG_ABS | RANGE
            out[bad] = float("nan")
            return out
        i = self.moved.index(name)
        return self.gathered[slot][:, i].reshape(n, self.plan.Fo, self.plan.A)

    def status(self, slot):
        return self.status_g[slot][:, :self.P_local].reshape(self.world * self.P_local)


def CopyEngineGather(plan, P_local, fields, group=None):
    """PushGather over the copy engines (one cudaMemcpyAsync per peer)."""
    return PushGather(plan, P_local, fields, group=group, transport="ce")


class PushGatherPass:
    """ONE pass over a rank's shard, cut into `pieces`: piece c is pushed into every peer's gathered
    buffer (same transports as `PushGather`) while piece c+1 is being computed, so that a pass costs
    max(kernels, gather) + one piece instead of kernels + gather.  For the strong-scaling
    configurations (BASELINE.json configs 3-5 split over the ranks), where a single all-gather at the
    end of the pass is fully exposed: config 5 on 8 GPUs gathers 3.4 GB of results to every rank, as
    long as its kernels take.

    Layout: `gathered` is (world, pieces, nfield, Pp, F, A) with Pp = ceil(P_local / pieces) -- one
    contiguous block per (rank, piece), written in place by kernel 3.  A ragged last piece leaves its
    unused rows unspecified.  `field(name)` is the (world, pieces, Pp, F, A) view; `valid_rows(c)` the
    number of profiles of piece c.  Every rank must pass the same P_local.  As in `PushGather`, `run()`
    starts with an entry barrier: consume a pass's results before calling `run()` again.
    """

    def __init__(self, plan, P_local, fields, pieces=4, group=None, transport="auto", nblocks=0):
        import os
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self._lib, self.lib = _lib, _lib.load()
        self.plan, self.fields = plan, tuple(fields)
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.P_local = int(P_local)
        pieces = max(1, min(int(pieces), self.P_local))
        self.Pp = -(-self.P_local // pieces)
        self.pieces = -(-self.P_local // self.Pp)
        F, A = plan.Fo, plan.A
        zero = () if plan.cloudy else ("tauliq", "tauice")
        self.moved = tuple(f for f in self.fields if f not in zero)
        self.zero_fields = tuple(f for f in self.fields if f in zero)
        nf = len(self.moved)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        self.Pp_pad = (self.Pp + 3) & ~3
        self.gathered = symm_mem.empty((self.world, self.pieces, nf, self.Pp, F, A), dtype=torch.float64, device=dev)
        self.hdl = symm_mem.rendezvous(self.gathered, self.group)
        self.status_g = symm_mem.empty((self.world, self.pieces, self.Pp_pad), dtype=torch.int32, device=dev)
        self.hdl_st = symm_mem.rendezvous(self.status_g, self.group)
        self.status_g.zero_()
        self.piece_bytes = nf * self.Pp * F * A * 8
        self.st_bytes = self.Pp_pad * 4
        mc = [int(getattr(h, "multicast_ptr", 0) or 0) for h in (self.hdl, self.hdl_st)]
        if transport == "auto":
            transport = "multicast" if (self.world > 4 and all(mc) and self.piece_bytes % 16 == 0) else "ce"
        if transport == "multicast" and not all(mc):
            raise RuntimeError("NVLink multicast is not available for this group")
        if transport in ("multicast", "peers") and self.piece_bytes % 16:
            raise RuntimeError("the push kernel moves 16-byte words; a piece has %d bytes" % self.piece_bytes)
        self.transport = transport
        self.nblocks = int(nblocks or os.environ.get("MWRT_PUSH_BLOCKS", 0) or 8)
        self.peers = [q for q in range(self.world) if q != self.rank]
        self.base = {q: int(self.hdl.buffer_ptrs[q]) for q in self.peers}
        self.base_st = {q: int(self.hdl_st.buffer_ptrs[q]) for q in self.peers}
        self.mc, self.mc_st = mc
        if transport == "ce":
            self.push = {q: torch.cuda.Stream() for q in self.peers}
        else:
            self.push = {None: torch.cuda.Stream(priority=-1)}
        self.comm = torch.cuda.Stream(priority=-1)
        self.computed = [torch.cuda.Event() for _ in range(self.pieces)]
        self.entered, self.free, self.landed = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        self.passes = 0

    def valid_rows(self, c):
        return min(self.P_local, (c + 1) * self.Pp) - c * self.Pp

    def _push(self, c):
        chk, lib = self._lib.check, self.lib
        off = (self.rank * self.pieces + c) * self.piece_bytes
        off_st = (self.rank * self.pieces + c) * self.st_bytes
        src = self.gathered[self.rank, c].data_ptr()
        src_st = self.status_g[self.rank, c].data_ptr()
        if self.transport == "ce":
            for q in self.peers:
                h = C.c_void_p(self.push[q].cuda_stream)
                chk(lib.mwrt_push_async(C.c_void_p(self.base[q] + off), C.c_void_p(src), self.piece_bytes, h), "push_async")
                chk(lib.mwrt_push_async(C.c_void_p(self.base_st[q] + off_st), C.c_void_p(src_st), self.st_bytes, h),
                    "push_async")
            return
        h = C.c_void_p(self.push[None].cuda_stream)
        if self.transport == "multicast":
            chk(lib.mwrt_push_kernel(C.c_void_p(src), self.piece_bytes, C.c_void_p(self.mc + off), None, 0,
                                     self.nblocks, h), "push_kernel")
            chk(lib.mwrt_push_kernel(C.c_void_p(src_st), self.st_bytes, C.c_void_p(self.mc_st + off_st), None, 0, 1, h),
                "push_kernel")
        else:
            n = len(self.peers)
            arr = (C.c_int64 * n)(*[self.base[q] + off for q in self.peers])
            chk(lib.mwrt_push_kernel(C.c_void_p(src), self.piece_bytes, None, arr, n, self.nblocks, h), "push_kernel")
            arr = (C.c_int64 * n)(*[self.base_st[q] + off_st for q in self.peers])
            chk(lib.mwrt_push_kernel(C.c_void_p(src_st), self.st_bytes, None, arr, n, 1, h), "push_kernel")

    def run(self, z, p, t, rh, denliq=None, denice=None, local_out=None):
        """`local_out`: dict of full-size (P_local, ...) tensors that stay on this rank (e.g. 'tau_layer'),
        sliced per piece."""
        main = torch.cuda.current_stream()
        self.entered.record(main)
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(self.entered)
            self.hdl_st.barrier(channel=0)          # every rank is past the consumers of the previous pass
            self.free.record(self.comm)
        if self.passes > 0:
            main.wait_event(self.landed)            # my pushes of the previous pass have left my slots
        for c in range(self.pieces):
            a = c * self.Pp
            n = self.valid_rows(c)
            own = self.gathered[self.rank, c]       # (nf, Pp, F, A)
            out = {k: own[i, :n] for i, k in enumerate(self.moved)}
            out["status"] = self.status_g[self.rank, c, :n]
            for k, v in (local_out or {}).items():
                out[k] = v[a:a + n]
            self.plan.forward(z[a:a + n], p[a:a + n], t[a:a + n], rh[a:a + n],
                              None if denliq is None else denliq[a:a + n],
                              None if denice is None else denice[a:a + n], fields=self.moved, out=out)
            self.computed[c].record(main)
            for st in self.push.values():
                st.wait_event(self.computed[c])
                if c == 0:
                    st.wait_event(self.free)
            self._push(c)
        with torch.cuda.stream(self.comm):
            for st in self.push.values():
                self.comm.wait_stream(st)
            self.hdl.barrier(channel=0)             # every rank's pieces have landed everywhere
            self.landed.record(self.comm)
        self.passes += 1

    def finish(self):
        torch.cuda.current_stream().wait_event(self.landed)

    def field(self, name):
        """(world, pieces, Pp, F, A) view of one gathered field (rank-major, then piece, then local index)."""
        if name in self.zero_fields:
            out = torch.zeros((self.world, self.pieces, self.Pp, self.plan.Fo, self.plan.A), dtype=torch.float64,
                              device=self.dev)
            out[(self.status_g[:, :, :self.Pp] & 11) != 0] = float("nan")
            return out
        return self.gathered[:, :, self.moved.index(name)]


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
