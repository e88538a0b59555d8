"""audio_processing mirror: the voice render / mix-down of Conductor::coordinate + Voice::process
(blast/src/audio_processing/engine.rs:46-81, 386-448) over the C ABI (blast_scene_*, blast_render), and the
Command-driven Conductor (engine.rs:36-248, commands.rs:86-234, blast_time.rs, processes.rs) over blast_conductor_*."""
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


# ---------------------------------------------------------------- Conductor (commands.rs / engine.rs names)
TM_PROCESS, TM_VOICE, TM_GROUP, TM_CONTEXT, TM_TBD = range(5)          # TempoMode (blast_time.rs:67-75)
TU_SAMPLES, TU_MILLIS, TU_BPM = range(3)                               # TempoUnit (blast_time.rs:77-82)
(CMD_LOAD, CMD_START, CMD_PAUSE, CMD_RESUME, CMD_STOP, CMD_UNLOAD, CMD_VELOCITY, CMD_GROUP, CMD_TC, CMD_SEQ,
 CMD_QUIT) = range(11)                                                 # Command (commands.rs:86-99)
IDX_TEMPO, IDX_VOICE, IDX_PROCESS, IDX_GROUP = range(4)                # Idx (commands.rs:163-169)


def tempo_repr(idx=0, owned=True, mode=TM_TBD, unit=TU_SAMPLES, interval=0.0) -> _lib.TempoRepr:
    """TempoRepr (commands.rs:187-234); the default is LoadArgs' no-`-t` value (commands.rs:197-206)."""
    return _lib.TempoRepr(idx, int(owned), mode, unit, interval)


def convert_interval(sample_rate: int, unit: int, interval: float) -> float:
    return _lib.load().blast_convert_interval(sample_rate, unit, interval)


class Cmd:
    """Builders for blast_command, one per Command variant.  Each returns (command, keepalive)."""

    @staticmethod
    def load(track_idx, tempo=None):
        return _lib.Command(kind=CMD_LOAD, idx=track_idx, tempo=tempo or tempo_repr())

    @staticmethod
    def transport(kind, idx, idx_kind=IDX_VOICE):
        return _lib.Command(kind=kind, idx_kind=idx_kind, idx=idx)

    @staticmethod
    def unload(idx):
        return _lib.Command(kind=CMD_UNLOAD, idx=idx)

    @staticmethod
    def velocity(idx, val):
        return _lib.Command(kind=CMD_VELOCITY, idx=idx, val=val)

    @staticmethod
    def tc(tempo):
        return _lib.Command(kind=CMD_TC, tempo=tempo)

    @staticmethod
    def group(tempo, members):
        """members: list of (voice_idx, update_tempo, [proc ids]) = GroupArgs.vs_fs_ps"""
        n = len(members)
        mv = (C.c_uint64 * max(1, n))(*[m[0] for m in members])
        mu = (C.c_uint8 * max(1, n))(*[int(m[1]) for m in members])
        mn = (C.c_uint32 * max(1, n))(*[len(m[2]) for m in members])
        flat = [p for m in members for p in m[2]]
        mp = (C.c_uint64 * max(1, len(flat)))(*flat)
        c = _lib.Command(kind=CMD_GROUP, tempo=tempo, n_members=n, member_voice=mv, member_update_tempo=mu,
                         member_n_procs=mn, member_proc_ids=mp)
        c._keep = (mv, mu, mn, mp)
        return c

    @staticmethod
    def seq(idx, tempo, period, steps, chance, rng_state, idx_kind=IDX_VOICE):
        n = len(steps)
        st = (C.c_float * max(1, n))(*steps)
        chn = (C.c_float * max(1, n))(*chance)
        c = _lib.Command(kind=CMD_SEQ, idx_kind=idx_kind, idx=idx, tempo=tempo, period=period, n_steps=n, steps=st,
                         chance=chn, rng_s0=rng_state[0], rng_s1=rng_state[1])
        c._keep = (st, chn)
        return c

    @staticmethod
    def quit():
        return _lib.Command(kind=CMD_QUIT)


class Conductor:
    """Conductor::{prepare, apply, coordinate} (engine.rs:36-248) on one GPU."""

    def __init__(self, ctx: Context, out_channels: int, sample_rate: int, tracks):
        self.ctx = ctx
        self.tracks = list(tracks)
        self.out_channels = out_channels
        t = (_lib.Track * max(1, len(tracks)))(*[x.c() for x in tracks])
        p = C.c_void_p()
        check(ctx.lib.blast_conductor_create(ctx.h, out_channels, sample_rate, t, len(tracks), C.byref(p)))
        self.h = p.value

    def close(self):
        if self.h and self.ctx.h:
            self.ctx.lib.blast_conductor_destroy(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def apply(self, cmd: _lib.Command):
        check(self.ctx.lib.blast_conductor_apply(self.ctx.h, self.h, C.byref(cmd)))

    def load(self, track_idx, tempo=None):
        self.apply(Cmd.load(track_idx, tempo))

    def start(self, idx, idx_kind=IDX_VOICE):
        self.apply(Cmd.transport(CMD_START, idx, idx_kind))

    def pause(self, idx, idx_kind=IDX_VOICE):
        self.apply(Cmd.transport(CMD_PAUSE, idx, idx_kind))

    def resume(self, idx, idx_kind=IDX_VOICE):
        self.apply(Cmd.transport(CMD_RESUME, idx, idx_kind))

    def stop(self, idx, idx_kind=IDX_VOICE):
        self.apply(Cmd.transport(CMD_STOP, idx, idx_kind))

    def unload(self, idx):
        self.apply(Cmd.unload(idx))

    def velocity(self, idx, val):
        self.apply(Cmd.velocity(idx, val))

    def tc(self, tempo):
        self.apply(Cmd.tc(tempo))

    def group(self, tempo, members):
        self.apply(Cmd.group(tempo, members))

    def seq(self, idx, tempo, period, steps, chance, rng_state, idx_kind=IDX_VOICE):
        self.apply(Cmd.seq(idx, tempo, period, steps, chance, rng_state, idx_kind))

    def set_shard(self, rank: int, world: int):
        check(self.ctx.lib.blast_conductor_set_shard(self.h, rank, world))

    def coordinate(self, frames: int) -> np.ndarray:
        bus = np.zeros(frames * self.out_channels, dtype=np.int16)
        check(self.ctx.lib.blast_conductor_coordinate(self.ctx.h, self.h, frames, bus.ctypes.data if bus.size else None))
        return bus

    def reserve(self, frames: int):
        """allocate now what a span of `frames` frames of the current scene needs (the spans that follow allocate nothing)"""
        check(self.ctx.lib.blast_conductor_reserve(self.ctx.h, self.h, frames))

    def render_partial_dev(self, frames: int, d_partial: int):
        check(self.ctx.lib.blast_conductor_render_dev(self.ctx.h, self.h, frames, d_partial))

    @staticmethod
    def _events(timeline):
        ev = (_lib.TimedCommand * max(1, len(timeline)))()
        for i, (frame, cmd) in enumerate(timeline):
            ev[i].frame = frame
            ev[i].cmd = cmd
        return ev

    def render_timeline(self, timeline, total_frames: int) -> np.ndarray:
        """timeline: [(frame, Command)] sorted by frame -> interleaved S16 bus of total_frames frames"""
        ev = self._events(timeline)
        bus = np.zeros(total_frames * self.out_channels, dtype=np.int16)
        check(self.ctx.lib.blast_conductor_render_timeline(self.ctx.h, self.h, ev, len(timeline), total_frames,
                                                           bus.ctypes.data if bus.size else None))
        return bus

    def render_timeline_partial_dev(self, timeline, total_frames: int, d_partial: int):
        ev = self._events(timeline)
        check(self.ctx.lib.blast_conductor_render_timeline_dev(self.ctx.h, self.h, ev, len(timeline), total_frames,
                                                               d_partial))

    def n_voices(self, group=-1):
        return self.ctx.lib.blast_conductor_n_voices(self.h, group)

    def n_groups(self):
        return self.ctx.lib.blast_conductor_n_groups(self.h)

    def get_voice(self, idx, group=-1) -> _lib.VoiceState:
        s = _lib.VoiceState()
        check(self.ctx.lib.blast_conductor_get_voice(self.h, group, idx, C.byref(s)))
        return s

    def set_voice(self, idx, group=-1, position=None, velocity=None, gain=None, active=None):
        def fp(x):
            return C.byref(C.c_float(x)) if x is not None else None
        act = C.byref(C.c_int(int(active))) if active is not None else None
        check(self.ctx.lib.blast_conductor_set_voice(self.h, group, idx, fp(position), fp(velocity), fp(gain), act))

    def clock(self):
        return self.ctx.lib.blast_conductor_clock(self.h)
