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
