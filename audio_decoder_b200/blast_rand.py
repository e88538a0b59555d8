"""blast_rand mirror: X128P parameter streams on the GPU (blast/src/audio_processing/blast_rand.rs:4-60)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .context import Context, DevBuf
from .errors import check


def seed_state(seed: int) -> tuple[int, int]:
    """X128P::new(seed) -> (s0, s1)"""
    st = _lib.X128PState()
    _lib.load().blast_x128p_seed(seed & (2**64 - 1), C.byref(st))
    return st.s0, st.s1


def advance(state, n_draws: int) -> tuple[int, int]:
    """jump one generator state by n_draws (host, GF(2) matrix power)"""
    a, b = _lib.X128PState(*state), _lib.X128PState()
    check(_lib.load().blast_x128p_advance(C.byref(a), n_draws, C.byref(b)))
    return b.s0, b.s1


class Streams:
    """n_streams generators; stream s starts s*stride draws into the sequence of `base` (seed or state)."""

    def __init__(self, ctx: Context, n_streams: int, stride: int, seed: int | None = None, state=None):
        self.ctx, self.n = ctx, n_streams
        base = _lib.X128PState(*(state if state is not None else seed_state(seed)))
        self.states = ctx.alloc(max(16, 16 * n_streams))
        check(ctx.lib.blast_x128p_jump_dev(ctx.h, C.byref(base), stride, n_streams, self.states.ptr))

    @classmethod
    def from_seeds(cls, ctx: Context, seeds) -> "Streams":
        """one independent generator per seed: stream i = X128P::new(seeds[i]) (blast_rand.rs:10-24)"""
        self = cls.__new__(cls)
        self.ctx, self.n = ctx, len(seeds)
        st = np.empty((max(1, self.n), 2), dtype=np.uint64)
        for i, s in enumerate(seeds):
            st[i] = seed_state(int(s))
        self.states = ctx.alloc(max(16, 16 * self.n)).upload(st)
        return self

    def get_states(self) -> np.ndarray:
        return self.states.download(np.uint64, 2 * self.n).reshape(self.n, 2)

    def fill_dev(self, draws: int, lo: int = 0, hi: int = 100, d_raw=None, d_ranged=None, d_checks=None):
        check(self.ctx.lib.blast_x128p_fill_dev(self.ctx.h, self.states.ptr, self.n, draws, lo, hi, d_raw, d_ranged,
                                                d_checks))

    def fill(self, draws: int, lo: int = 0, hi: int = 100, raw=True, ranged=True, checks=True):
        """-> dict of host arrays (raw uint64 [n,draws], ranged int64 [n,draws], checks uint64 [n,4])"""
        bufs = {}
        if raw:
            bufs["raw"] = self.ctx.alloc(max(16, 8 * self.n * draws))
        if ranged:
            bufs["ranged"] = self.ctx.alloc(max(16, 8 * self.n * draws))
        if checks:
            bufs["checks"] = self.ctx.alloc(max(16, 32 * self.n))
        self.fill_dev(draws, lo, hi, bufs["raw"].ptr if raw else None, bufs["ranged"].ptr if ranged else None,
                      bufs["checks"].ptr if checks else None)
        out = {}
        if raw:
            out["raw"] = bufs["raw"].download(np.uint64, self.n * draws).reshape(self.n, draws)
        if ranged:
            out["ranged"] = bufs["ranged"].download(np.int64, self.n * draws).reshape(self.n, draws)
        if checks:
            out["checks"] = bufs["checks"].download(np.uint64, self.n * 4).reshape(self.n, 4)
        return out


def fill(ctx: Context, seed: int, stride: int, n_streams: int, draws: int, lo: int = 0, hi: int = 100,
         raw=True, ranged=True, checks=True):
    """blast_x128p_fill host one-shot"""
    r = np.empty((n_streams, draws), dtype=np.uint64) if raw else None
    g = np.empty((n_streams, draws), dtype=np.int64) if ranged else None
    c = np.empty((n_streams, 4), dtype=np.uint64) if checks else None
    check(ctx.lib.blast_x128p_fill(ctx.h, seed & (2**64 - 1), stride, n_streams, draws, lo, hi,
                                   r.ctypes.data if raw else None, g.ctypes.data if ranged else None,
                                   c.ctypes.data if checks else None))
    return r, g, c
