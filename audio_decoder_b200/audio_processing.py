"""audio_processing mirror: the voice render / mix-down of Conductor::coordinate + Voice::process
(blast/src/audio_processing/engine.rs:46-81, 386-448) over the C ABI (blast_scene_*, blast_render)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .context import Context, DevBuf
from .errors import check


@dataclass
class Track:
    """AudioFile samples resident in HBM (what Voice::new clones, engine.rs:298-316)."""
    buf: DevBuf
    n_samples: int
    num_channels: int
    sample_rate: int = 44100

    @classmethod
    def from_host(cls, ctx: Context, samples: np.ndarray, num_channels: int, sample_rate: int = 44100) -> "Track":
        s = np.ascontiguousarray(samples, dtype=np.int16)
        return cls(ctx.to_device(s) if s.size else ctx.alloc(16), s.size, num_channels, sample_rate)

    def c(self) -> _lib.Track:
        return _lib.Track(self.buf.ptr, self.n_samples, self.num_channels, self.sample_rate)


@dataclass
class VoiceParams:
    """VoiceState (engine.rs:279-286) minus the tempo; defaults as Voice::new."""
    track: int
    active: bool = False
    position: float = 0.0
    velocity: float = 1.0
    gain: float = 1.0

    def c(self) -> _lib.Voice:
        return _lib.Voice(self.track, int(self.active), self.position, self.velocity, self.gain, 0)


class Scene:
    """A fixed set of voices on one GPU; render() == `frames` iterations of coordinate()'s frame loop."""

    def __init__(self, ctx: Context, tracks, voices, out_channels: int):
        self.ctx = ctx
        self.tracks = list(tracks)
        self.n_voices = len(voices)
        self.out_channels = out_channels
        t = (_lib.Track * max(1, len(tracks)))(*[x.c() for x in tracks])
        v = (_lib.Voice * max(1, len(voices)))(*[x.c() for x in voices])
        p = C.c_void_p()
        check(ctx.lib.blast_scene_create(ctx.h, t, len(tracks), v, len(voices), out_channels, C.byref(p)))
        self.h = p.value

    def close(self):
        if self.h and self.ctx.h:
            self.ctx.lib.blast_scene_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_voices(self, voices):
        v = (_lib.Voice * max(1, len(voices)))(*[x.c() for x in voices])
        check(self.ctx.lib.blast_scene_set_voices(self.ctx.h, self.h, v, len(voices)))

    def restore_dev(self):
        """async rewind to the voice table of the last create / set_voices"""
        check(self.ctx.lib.blast_scene_restore_dev(self.ctx.h, self.h))

    def voices(self):
        v = (_lib.Voice * max(1, self.n_voices))()
        check(self.ctx.lib.blast_scene_get_voices(self.ctx.h, self.h, v, self.n_voices))
        return [VoiceParams(x.track, bool(x.active), x.position, x.velocity, x.gain) for x in v[:self.n_voices]]

    def render_partial_dev(self, frames: int, d_partial: int):
        """async: int32 partial bus [frames * out_channels] at device address d_partial"""
        check(self.ctx.lib.blast_scene_render_dev(self.ctx.h, self.h, frames, d_partial))

    def check(self):
        check(self.ctx.lib.blast_scene_check(self.ctx.h, self.h))

    def render(self, frames: int) -> np.ndarray:
        """-> interleaved S16 bus (host), voices advanced"""
        n = frames * self.out_channels
        part = self.ctx.alloc(max(16, 4 * n))
        bus = self.ctx.alloc(max(16, 2 * n))
        self.render_partial_dev(frames, part.ptr)
        check(self.ctx.lib.blast_bus_finalize_dev(self.ctx.h, part.ptr, bus.ptr, n))
        self.check()
        return bus.download(np.int16, n)


def render(ctx: Context, tracks, voices, out_channels: int, frames: int):
    """blast_render one-shot -> (bus int16 [frames*out_channels], voices after)"""
    t = (_lib.Track * max(1, len(tracks)))(*[x.c() for x in tracks])
    v = (_lib.Voice * max(1, len(voices)))(*[x.c() for x in voices])
    after = (_lib.Voice * max(1, len(voices)))()
    bus = np.zeros(frames * out_channels, dtype=np.int16)
    check(ctx.lib.blast_render(ctx.h, t, len(tracks), v, len(voices), out_channels, frames,
                               bus.ctypes.data if bus.size else None, after))
    return bus, [VoiceParams(x.track, bool(x.active), x.position, x.velocity, x.gain) for x in after[:len(voices)]]


def finalize_bus(ctx: Context, d_partial: int, d_bus: int, n_slots: int):
    check(ctx.lib.blast_bus_finalize_dev(ctx.h, d_partial, d_bus, n_slots))
