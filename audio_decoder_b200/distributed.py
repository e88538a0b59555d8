"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

The hot path shards naturally (SURVEY.md §8 e): files, voices and RNG streams are independent, item i
goes to rank i mod world and no data-path collective is needed — except for the render, where every
rank mixes its own voices into an int32 partial bus and ONE all-reduce (sum) produces the full bus.
That is exact: the reference's i16 wrapping accumulate (engine.rs:441) is addition mod 2^16, so
partial sums may be formed in any order / on any GPU in int32 and wrapped at the end
(4,096 voices x 32,768 < 2^31, and int32 wrap-around is harmless mod 2^16 anyway).  NCCL has no
int16 type, so the bus travels as int32.
"""
from __future__ import annotations

import numpy as np


def shard(n_items: int, rank: int, world: int) -> range:
    """item i -> rank i mod world"""
    return range(rank, n_items, world)


def shard_counts(n_items: int, world: int):
    return [len(shard(n_items, r, world)) for r in range(world)]


def all_reduce_partial_bus(partial, group=None):
    """in-place sum of the int32 partial buses of all ranks (torch tensor, CPU/gloo or CUDA/NCCL)"""
    import torch
    import torch.distributed as dist
    assert partial.dtype == torch.int32
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial


def wrap_i16(partial):
    """int32 partial sums -> the S16 bus (low 16 bits); numpy or torch, host-side twin of blast_bus_finalize_dev"""
    if isinstance(partial, np.ndarray):
        return partial.astype(np.int32).astype(np.int16)
    import torch
    return partial.to(torch.int16)


class ShardedScene:
    """Voices sharded over the ranks of a process group; render() returns the full S16 bus on every rank.

    ctx must have been created on torch's current CUDA stream (Context(device, stream=...)) so that the
    render kernels, the NCCL all-reduce and the finalize kernel are ordered on one stream.
    """

    def __init__(self, ctx, tracks, voices, out_channels: int, rank: int, world: int, group=None):
        from . import audio_processing as ap
        self.ctx, self.group, self.rank, self.world = ctx, group, rank, world
        self.out_channels = out_channels
        mine = list(shard(len(voices), rank, world))
        self.voice_ids = mine
        self.scene = ap.Scene(ctx, tracks, [voices[i] for i in mine], out_channels)

    def render(self, frames: int):
        import torch
        from . import audio_processing as ap
        n = frames * self.out_channels
        dev = torch.device("cuda", self.ctx.device)
        part = torch.empty(n, dtype=torch.int32, device=dev)
        bus = torch.empty(n, dtype=torch.int16, device=dev)
        self.scene.render_partial_dev(frames, part.data_ptr())
        all_reduce_partial_bus(part, self.group)
        ap.finalize_bus(self.ctx, part.data_ptr(), bus.data_ptr(), n)
        self.scene.check()
        return bus


# ---------------------------------------------------------------------------------- MPEG scan over byte ranges
MPEG_RANGE_ALIGN = 32768          # every range but the last is a multiple of the scan's span size
MPEG_HALO = 16                    # bytes of the following range a rank must also hold (look-ahead + header bytes)
MPEG_HDR_BINS = 1 << 21


def mpeg_plan_ranges(total_len: int, world: int):
    """-> [(start, own_len, halo_len)] per rank: contiguous ranges, all but the last a multiple of 32 KiB"""
    spans = (total_len + MPEG_RANGE_ALIGN - 1) // MPEG_RANGE_ALIGN
    per = (spans + world - 1) // world
    out = []
    for r in range(world):
        a = min(total_len, r * per * MPEG_RANGE_ALIGN)
        b = min(total_len, (r + 1) * per * MPEG_RANGE_ALIGN)
        halo = min(MPEG_HALO, total_len - b)
        # a range that ends the stream carries no halo; ranges after it are empty
        out.append((a, b - a, halo if b < total_len else 0))
    return out


def mpeg_fold_aggs(aggs):
    """aggs[r] = (exit_state[4], count[4]) of range r -> [(entry_state, candidates_before)] per rank, total.
    The greedy scan is a 4-state machine (header bytes still to skip); range 0 starts in state 0."""
    state, before, out = 0, 0, []
    for exit_state, count in aggs:
        out.append((state, before))
        before += int(count[state])
        state = int(exit_state[state])
    return out, before


def mpeg_exchange_aggs(agg, group=None, device=None):
    """all-gather of the 8 numbers of every rank's range aggregate (the scan's one exchange step)"""
    import torch
    import torch.distributed as dist
    mine = torch.tensor(list(agg[0]) + list(agg[1]), dtype=torch.int64, device=device)
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return [agg]
    world = dist.get_world_size(group)
    got = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(got, mine, group=group)
    return [([int(x) for x in g[:4].tolist()], [int(x) for x in g[4:].tolist()]) for g in got]


class ShardedMpegIndex:
    """mpeg::parse's frame-offset index for ONE stream cut into byte ranges over the ranks of a process group.

    Every rank holds its range (+ 16 halo bytes) in HBM.  Collectives: one all-gather of 8 int64 per rank (range
    aggregates), one all-reduce(sum) of the 2^21-bin header histogram (8 MB, int32), and with reference_compat one
    all-reduce(min) of the first-position table (16 MB, int64).  The index stays sharded: each rank returns the
    offsets that fall into its range (they are global file positions, already in order across ranks).
    """

    def __init__(self, ctx, rank: int, world: int, group=None):
        self.ctx, self.rank, self.world, self.group = ctx, rank, world, group

    def run(self, d_bytes: int, total_len: int, reference_compat: bool = True):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import file_parsing as fp
        from .errors import check
        ctx = self.ctx
        dev = torch.device("cuda", ctx.device)
        start, own, halo = mpeg_plan_ranges(total_len, self.world)[self.rank]
        agg = fp.mpeg.shard_walk_dev(ctx, d_bytes, own, halo)
        aggs = mpeg_exchange_aggs(agg, self.group, dev)
        folded, n_cand_total = mpeg_fold_aggs(aggs)
        entry, _before = folded[self.rank]
        count = agg[1][entry]
        d_pos, d_hdr = fp.mpeg.shard_emit_dev(ctx, d_bytes, own, halo, entry, start, count)
        hist = torch.zeros(MPEG_HDR_BINS, dtype=torch.int32, device=dev)
        check(ctx.lib.blast_mpeg_hist_dev(ctx.h, d_hdr.ptr, count, hist.data_ptr()))
        multi = dist.is_initialized() and dist.get_world_size(self.group) > 1
        if multi:
            dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=self.group)
        ref = C.c_uint32()
        check(ctx.lib.blast_mpeg_pick_ref_dev(ctx.h, hist.data_ptr(), C.byref(ref)))
        first_ptr = None
        if reference_compat:
            # all-ones = "no position yet"; int64 max is the same bit pattern minus the sign bit, so min() over
            # int64 needs non-negative values: use 2^63-1 as the sentinel (positions are far below it)
            first = torch.full((MPEG_HDR_BINS,), (1 << 63) - 1, dtype=torch.int64, device=dev)
            check(ctx.lib.blast_mpeg_first_pos_dev(ctx.h, d_pos.ptr, d_hdr.ptr, count, ref.value, first.data_ptr()))
            if multi:
                dist.all_reduce(first, op=dist.ReduceOp.MIN, group=self.group)
            first_ptr = first.data_ptr()
        n = C.c_uint64()
        check(ctx.lib.blast_mpeg_classify_dev(ctx.h, d_pos.ptr, d_hdr.ptr, count, ref.value, first_ptr, total_len, None, 0,
                                              C.byref(n)))
        d_off = ctx.alloc(max(16, 8 * n.value))
        check(ctx.lib.blast_mpeg_classify_dev(ctx.h, d_pos.ptr, d_hdr.ptr, count, ref.value, first_ptr, total_len,
                                              d_off.ptr, n.value, C.byref(n)))
        return dict(d_offsets=d_off, n_offsets=n.value, ref_header=ref.value, n_candidates=n_cand_total,
                    n_candidates_local=count, range=(start, own, halo), d_pos=d_pos, d_hdr=d_hdr)


# ---------------------------------------------------------------------------------- the mix reduction over peer memory
class PeerBus:
    """The render's one exchange step without a collective library: every rank's int32 partial bus and flag block are
    mapped into every peer's address space (CUDA IPC over NVLink / NVSwitch); each rank reduces ITS 1/N slice of all
    buses, wraps it to S16 and stores it straight into the root's bus — one kernel per rank
    (blast_bus_reduce_peers_dev).  torch.distributed only carries the 64-byte IPC handles at set-up time.

    Per step, on every rank:      pb.wait_ack(); <render into pb.part.ptr>; pb.reduce()      (the bus: pb.bus on the root)
    Flag block of a rank (uint32 step counters): ready[w] at byte 0, ack[w] at byte 256, done[w] at byte 512.
    """
    READY, ACK, DONE = 0, 256, 512

    def __init__(self, ctx, n_slots: int, rank: int, world: int, group=None, root: int = 0, mode: str = "root"):
        """mode "root":    the root pulls and reduces the whole of every peer's bus (the other ranks only signal and
                           never wait for each other: no lock step; the root's NVLink ingress carries N-1 buses);
           mode "scatter": every rank reduces its 1/N slice and stores it into the root's bus (1/N of the traffic per
                           GPU, but every rank waits for every other one each step)."""
        import ctypes as C
        import torch.distributed as dist
        from .errors import check
        assert world <= 16 and mode in ("root", "scatter")
        self.mode = mode
        self.ctx, self.rank, self.world, self.root, self.n_slots = ctx, rank, world, root, n_slots
        self.part = ctx.alloc(max(256, 4 * n_slots))
        self.bus = ctx.alloc(max(256, 2 * n_slots))          # the S16 bus (complete on the root only)
        self.flags = ctx.alloc(1024)
        self.flags.zero()
        self.part.zero()
        ctx.sync()
        self.step = 0
        self._opened = []
        # my slice of the bus: multiples of 8 slots so that int32 / int16 vector accesses stay aligned
        per = ((n_slots + world - 1) // world + 7) // 8 * 8
        self.slot0 = min(n_slots, rank * per)
        self.slice_len = min(n_slots, (rank + 1) * per) - self.slot0
        if world == 1:
            self.peer_parts, self.peer_flags, self.root_bus = [], [], self.bus.ptr
            return

        def export(ptr):
            h = (C.c_uint8 * 64)()
            check(ctx.lib.blast_ipc_export(ctx.h, ptr, h))
            return bytes(h)

        def open_(handle):
            p = C.c_void_p()
            check(ctx.lib.blast_ipc_open(ctx.h, (C.c_uint8 * 64).from_buffer_copy(handle), C.byref(p)))
            self._opened.append(p.value)
            return p.value

        # every rank reaches every collective below even if its own IPC calls fail, and all ranks then agree on the
        # outcome: either everybody has the mapping or everybody raises (the caller may then choose the NCCL variant)
        err = None
        try:
            mine = (export(self.part.ptr), export(self.flags.ptr), export(self.bus.ptr))
        except Exception as e:                                         # noqa: BLE001
            mine, err = None, f"rank {rank}: {e}"
        everyone = [None] * world
        dist.all_gather_object(everyone, mine, group=group)
        self.peers = [r for r in range(world) if r != rank]
        if err is None and all(x is not None for x in everyone):
            try:
                self.peer_parts = [open_(everyone[r][0]) for r in self.peers]
                self.peer_flags = [open_(everyone[r][1]) for r in self.peers]
                self.root_bus = self.bus.ptr if rank == root else open_(everyone[root][2])
                self.root_flags = self.flags.ptr if rank == root else self.peer_flags[self.peers.index(root)]
            except Exception as e:                                     # noqa: BLE001
                err = f"rank {rank}: {e}"
        elif err is None:
            err = "a peer could not export its buffers"
        errs = [None] * world
        dist.all_gather_object(errs, err, group=group)
        bad = [e for e in errs if e]
        if bad:
            for p in self._opened:
                ctx.lib.blast_ipc_close(ctx.h, p)
            self._opened = []
            raise RuntimeError("peer memory (CUDA IPC) is not available on this box: " + "; ".join(bad))

    def _ptrs(self, values):
        import ctypes as C
        return (C.c_void_p * max(1, len(values)))(*values), len(values)

    def wait_ack(self):
        """before overwriting the partial bus again: the stream waits until every rank is done reading the previous one"""
        from .errors import check
        if self.world > 1 and self.step > 0:
            if self.mode == "scatter":
                check(self.ctx.lib.blast_peer_wait_dev(self.ctx.h, self.flags.ptr + self.ACK, self.world, self.step))
            elif self.rank != self.root:
                check(self.ctx.lib.blast_peer_wait_dev(self.ctx.h, self.flags.ptr + self.ACK + 4 * self.root, 1, self.step))

    def reduce(self):
        """after the render: publish the partial bus, reduce + wrap this rank's slice into the root's bus; on the root the
        stream then waits until every slice is in place (the bus is complete for whatever is enqueued next)"""
        from .errors import check
        self.step += 1
        L, ctx, w = self.ctx.lib, self.ctx, self.world
        if w == 1:
            check(L.blast_bus_finalize_dev(ctx.h, self.part.ptr, self.bus.ptr, self.n_slots))
            return
        me = 4 * self.rank
        if self.mode == "root":
            mine, n = self._ptrs([self.root_flags + self.READY + me])
            check(L.blast_peer_signal_dev(ctx.h, mine, n, self.step))
            if self.rank == self.root:
                parts, n_parts = self._ptrs([self.part.ptr] + self.peer_parts)
                after, n_after = self._ptrs([f + self.ACK + me for f in self.peer_flags])
                check(L.blast_bus_reduce_peers_dev(ctx.h, parts, n_parts, self.flags.ptr + self.READY, w, self.step,
                                                   self.bus.ptr, 0, self.n_slots, after, n_after))
            return
        ready, n = self._ptrs([f + self.READY + me for f in self.peer_flags] + [self.flags.ptr + self.READY + me])
        check(L.blast_peer_signal_dev(ctx.h, ready, n, self.step))
        parts, n_parts = self._ptrs([self.part.ptr] + self.peer_parts)
        after, n_after = self._ptrs([f + self.ACK + me for f in self.peer_flags] + [self.flags.ptr + self.ACK + me,
                                                                                  self.root_flags + self.DONE + me])
        check(L.blast_bus_reduce_peers_dev(ctx.h, parts, n_parts, self.flags.ptr + self.READY, w, self.step, self.root_bus,
                                           self.slot0, self.slice_len, after, n_after))
        if self.rank == self.root:
            check(L.blast_peer_wait_dev(ctx.h, self.flags.ptr + self.DONE, w, self.step))

    def close(self):
        self.ctx.sync()
        for p in self._opened:
            self.ctx.lib.blast_ipc_close(self.ctx.h, p)
        self._opened = []


class ShardedConductor:
    """The Command-driven Conductor (engine.rs:36-248) over the ranks of a process group: every rank applies every
    command and tracks every tempo, renders the voices whose load number is congruent to its rank, and the partial
    buses are reduced + finalized over peer memory (PeerBus).  coordinate() returns the S16 bus on the root rank
    (None elsewhere).  max_frames = the longest coordinate() span that will be asked for."""

    def __init__(self, ctx, out_channels: int, sample_rate: int, tracks, max_frames: int, rank: int, world: int,
                 group=None, root: int = 0, mode: str = "root"):
        from . import audio_processing as ap
        self.ctx, self.rank, self.world, self.root = ctx, rank, world, root
        self.out_channels, self.max_frames = out_channels, max_frames
        self.conductor = ap.Conductor(ctx, out_channels, sample_rate, tracks)
        self.conductor.set_shard(rank, world)
        self.peer = PeerBus(ctx, max_frames * out_channels, rank, world, group=group, root=root, mode=mode)

    def __getattr__(self, name):                      # load / start / stop / velocity / seq / group / tc / apply / set_voice ...
        return getattr(self.conductor, name)

    def coordinate(self, frames: int):
        assert frames <= self.max_frames
        n = frames * self.out_channels
        self.peer.wait_ack()
        # the reduction always covers the whole mapped bus: clear the tail this span does not write
        if n < self.peer.n_slots:
            from .errors import check
            check(self.ctx.lib.blast_memset_dev(self.ctx.h, self.peer.part.ptr + 4 * n, 0, 4 * (self.peer.n_slots - n)))
        self.conductor.render_partial_dev(frames, self.peer.part.ptr)
        self.peer.reduce()
        if self.rank == self.root:
            return self.peer.bus.download(np.int16, n)
        self.ctx.sync()
        return None

    def close(self):
        self.peer.close()
        self.conductor.close()
