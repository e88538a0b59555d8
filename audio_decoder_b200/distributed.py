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
    """The render's one exchange step without a collective library (include/blast_cuda.h, blast_peer_bus_*): every
    rank's window — int32 partial bus, S16 bus, flag tables — is mapped into every peer's address space (CUDA IPC over
    NVLink / NVSwitch); the bus is cut into tiles, tile t is reduced by rank t mod world — by two small kernels after the
    render, or (fused=True) inside the render kernel as the tiles complete.  This class only creates the window and
    carries the 64-byte handles (torch.distributed, set-up time); the protocol itself lives in the library.

    Per step, on every rank:      pb.render_reduce(scene, frames)     before the bus is read:   pb.wait();  bus = pb.bus_ptr
    or (Conductor spans):         pb.begin(); <render into pb.part_ptr>; pb.reduce(n_slots)
    The exchange runs on the peer bus's own stream, beside what is enqueued next; wait() joins it.
    """

    def __init__(self, ctx, n_slots: int, rank: int, world: int, group=None, root: int = 0, fused: bool = False):
        """fused: render_reduce() does the exchange inside the render kernel instead of as two kernels after it"""
        import ctypes as C
        from .errors import check
        self.ctx, self.rank, self.world, self.root, self.n_slots = ctx, rank, world, root, n_slots
        self.h = None
        p = C.c_void_p()
        err = None
        try:
            check(ctx.lib.blast_peer_bus_create(ctx.h, n_slots, rank, world, root, C.byref(p)))
            self.h = p.value
        except Exception as e:                                         # noqa: BLE001
            err = f"rank {rank}: {e}"
        if world > 1:
            import torch.distributed as dist
            # every rank reaches every collective below even if its own calls fail, and all ranks agree on the
            # outcome: either everybody has the mapping or everybody raises (the caller may then choose NCCL)
            mine = None
            if err is None:
                try:
                    hb = (C.c_uint8 * 64)()
                    check(ctx.lib.blast_peer_bus_export(ctx.h, self.h, hb))
                    mine = bytes(hb)
                except Exception as e:                                 # noqa: BLE001
                    err = f"rank {rank}: {e}"
            everyone = [None] * world
            dist.all_gather_object(everyone, mine, group=group)
            if err is None and all(x is not None for x in everyone):
                try:
                    blob = (C.c_uint8 * (64 * world)).from_buffer_copy(b"".join(everyone))
                    check(ctx.lib.blast_peer_bus_connect_ipc(ctx.h, self.h, blob))
                except Exception as e:                                 # noqa: BLE001
                    err = f"rank {rank}: {e}"
            elif err is None:
                err = f"rank {rank}: a peer could not export its window"
            errs = [None] * world
            dist.all_gather_object(errs, err, group=group)
            bad = [e for e in errs if e]
            if bad:
                self.close()
                raise RuntimeError("peer memory (CUDA IPC) is not available on this box: " + "; ".join(bad))
        elif err:
            raise RuntimeError(err)
        check(ctx.lib.blast_peer_bus_set_fused(self.h, int(fused)))
        self.part_ptr = ctx.lib.blast_peer_bus_partial(self.h)
        self.bus_ptr = ctx.lib.blast_peer_bus_bus(self.h)

    def render_reduce(self, scene, frames: int):
        """async: render the scene's voices of this rank and reduce the bus tile by tile inside the render kernel"""
        from .errors import check
        check(self.ctx.lib.blast_scene_render_reduce_dev(self.ctx.h, scene.h, frames, self.h))

    def begin(self):
        """async: before this rank overwrites its partial bus outside render_reduce (every rank is done reading it)"""
        from .errors import check
        check(self.ctx.lib.blast_peer_bus_begin_dev(self.ctx.h, self.h))

    def reduce(self, n_slots: int | None = None):
        """async: publish + reduce the partial bus filled by earlier stream work (the first n_slots slots)"""
        from .errors import check
        check(self.ctx.lib.blast_peer_bus_reduce_dev(self.ctx.h, self.h, self.n_slots if n_slots is None else n_slots))

    def wait(self):
        """async: the stream joins the exchange of the current step (on the root: every rank's tiles are in the bus)"""
        from .errors import check
        check(self.ctx.lib.blast_peer_bus_wait_dev(self.ctx.h, self.h))

    def check(self):
        """synchronises; raises BlastError(ERR_TIMEOUT) if a device-side wait for a peer gave up"""
        from .errors import check
        check(self.ctx.lib.blast_peer_bus_check(self.ctx.h, self.h))

    def download_bus(self, n_slots: int | None = None):
        """root: wait + copy the S16 bus to the host"""
        import ctypes as C
        from .errors import check
        n = self.n_slots if n_slots is None else n_slots
        self.wait()
        out = np.empty(n, dtype=np.int16)
        check(self.ctx.lib.blast_memcpy_d2h(self.ctx.h, out.ctypes.data, self.bus_ptr, out.nbytes))
        self.check()
        return out

    def close(self):
        if self.h and self.ctx.h:
            self.ctx.lib.blast_peer_bus_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedConductor:
    """The Command-driven Conductor (engine.rs:36-248) over the ranks of a process group: every rank applies every
    command and tracks every tempo, renders the voices whose load number is congruent to its rank, and the partial
    buses are reduced + finalized over peer memory (PeerBus).  coordinate() returns the S16 bus on the root rank
    (None elsewhere).  max_frames = the longest coordinate() span that will be asked for."""

    def __init__(self, ctx, out_channels: int, sample_rate: int, tracks, max_frames: int, rank: int, world: int,
                 group=None, root: int = 0):
        from . import audio_processing as ap
        self.ctx, self.rank, self.world, self.root = ctx, rank, world, root
        self.out_channels, self.max_frames = out_channels, max_frames
        self.conductor = ap.Conductor(ctx, out_channels, sample_rate, tracks)
        self.conductor.set_shard(rank, world)
        self.peer = PeerBus(ctx, max_frames * out_channels, rank, world, group=group, root=root)

    def __getattr__(self, name):                      # load / start / stop / velocity / seq / group / tc / apply / set_voice ...
        return getattr(self.conductor, name)

    def coordinate(self, frames: int):
        assert frames <= self.max_frames
        n = frames * self.out_channels
        self.peer.begin()
        failure = None
        try:
            self.conductor.render_partial_dev(frames, self.peer.part_ptr)
        except Exception as e:                        # noqa: BLE001
            # a rank-local failure (capacity overflow, allocation) must not leave the peers waiting for this rank's
            # tiles: the step is published all the same (its bus is void), then the error is raised here
            failure = e
        self.peer.reduce(n)
        if self.rank == self.root:
            bus = self.peer.download_bus(n)
        else:
            self.peer.check()
            bus = None
        if failure is not None:
            raise failure
        return bus

    def close(self):
        self.peer.close()
        self.conductor.close()
