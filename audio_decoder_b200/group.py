"""blast_group mirror (include/blast_cuda.h, "several GPUs, one host process"): what a single-process host like the
reference's main() (blast/src/main.rs:13-128) uses to drive every GPU of a box — decode sharded by file, render sharded
by track with the bus reduced over peer memory inside the render kernel, RNG sharded by stream, MPEG by byte range.
All of it is the library's C++; this file only marshals arguments."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .errors import check


class Group:
    def __init__(self, devices, fused: bool = False):
        """fused: render() does the bus exchange inside the render kernel instead of as two kernels after it"""
        self.lib = _lib.load()
        ids = (C.c_int * len(devices))(*devices)
        p = C.c_void_p()
        check(self.lib.blast_group_create(C.byref(p), ids, len(devices)))
        self.h = p.value
        self.n = len(devices)
        check(self.lib.blast_group_set_fused(self.h, int(fused)))

    def close(self):
        if getattr(self, "h", None):
            self.lib.blast_group_destroy(self.h)
        self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def member_context(self, member: int):
        """member's blast_ctx as a (non-owning) Context: uploads / downloads on that member's GPU"""
        from .context import BorrowedContext
        return BorrowedContext(self.lib, self.lib.blast_group_ctx(self.h, member))

    # ---- main.rs:18-89: decode every asset; file i lands on member i mod n
    def decode_batch(self, images, descs, to_host: bool = True):
        """-> (host sample arrays or None, ctypes array of blast_track living on the members)"""
        n = len(images)
        keep = [np.ascontiguousarray(np.frombuffer(im, dtype=np.uint8) if not isinstance(im, np.ndarray) else im) for im in images]
        files = (C.c_void_p * max(1, n))(*[k.ctypes.data for k in keep])
        lens = (C.c_size_t * max(1, n))(*[k.size for k in keep])
        dd = (_lib.PcmDesc * max(1, n))(*descs)
        outs = [np.empty(self.lib.blast_pcm_out_len(C.byref(d)), dtype=np.int16) for d in descs] if to_host else None
        ho = (C.c_void_p * max(1, n))(*[o.ctypes.data for o in outs]) if to_host else None
        tracks = (_lib.Track * max(1, n))()
        check(self.lib.blast_group_pcm_decode_batch(self.h, n, files, lens, dd, ho, tracks))
        return outs, tracks

    def free_tracks(self):
        check(self.lib.blast_group_free_tracks(self.h))

    def render(self, tracks, n_tracks: int, voices, out_channels: int, frames: int) -> np.ndarray:
        v = (_lib.Voice * max(1, len(voices)))(*[x.c() for x in voices])
        bus = np.zeros(frames * out_channels, dtype=np.int16)
        check(self.lib.blast_group_render(self.h, tracks, n_tracks, v, len(voices), out_channels, frames,
                                          bus.ctypes.data if bus.size else None))
        return bus

    def x128p_fill(self, seed: int, stride: int, n_streams: int, draws: int, lower: int = 0, upper: int = 100):
        raw = np.empty((n_streams, draws), dtype=np.uint64)
        ranged = np.empty((n_streams, draws), dtype=np.int64)
        checks = np.empty((n_streams, 4), dtype=np.uint64)
        check(self.lib.blast_group_x128p_fill(self.h, seed, stride, n_streams, draws, lower, upper, raw.ctypes.data,
                                              ranged.ctypes.data, checks.ctypes.data))
        return raw, ranged, checks

    def mpeg_index(self, stream: np.ndarray, reference_compat: bool = True):
        s = np.ascontiguousarray(stream, dtype=np.uint8)
        n, ref, ncand = C.c_uint64(), C.c_uint32(), C.c_uint64()
        check(self.lib.blast_group_mpeg_index(self.h, s.ctypes.data, s.size, int(reference_compat), None, 0, C.byref(n),
                                              C.byref(ref), C.byref(ncand)))
        off = np.empty(n.value, dtype=np.uint64)
        check(self.lib.blast_group_mpeg_index(self.h, s.ctypes.data, s.size, int(reference_compat), off.ctypes.data, off.size,
                                              C.byref(n), C.byref(ref), C.byref(ncand)))
        return dict(offsets=off, ref_header=ref.value, n_candidates=ncand.value)


class GroupConductor:
    """Conductor::{prepare, apply, coordinate} (engine.rs:36-248) over a Group; same methods as audio_processing.Conductor."""

    def __init__(self, group: Group, out_channels: int, sample_rate: int, tracks, n_tracks: int):
        self.group, self.out_channels = group, out_channels
        p = C.c_void_p()
        check(group.lib.blast_group_conductor_create(group.h, out_channels, sample_rate, tracks, n_tracks, C.byref(p)))
        self.h = p.value

    def close(self):
        if self.h and self.group.h:
            self.group.lib.blast_group_conductor_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def apply(self, cmd):
        check(self.group.lib.blast_group_conductor_apply(self.h, C.byref(cmd)))

    def __getattr__(self, name):
        # load / start / stop / velocity / seq / group / tc ... : audio_processing.Cmd builds the command
        from .audio_processing import Cmd, CMD_START, CMD_PAUSE, CMD_RESUME, CMD_STOP, IDX_VOICE
        kinds = {"start": CMD_START, "pause": CMD_PAUSE, "resume": CMD_RESUME, "stop": CMD_STOP}
        if name in kinds:
            return lambda idx, idx_kind=IDX_VOICE: self.apply(Cmd.transport(kinds[name], idx, idx_kind))
        if name in ("load", "unload", "velocity", "tc", "group", "seq", "quit"):
            return lambda *a, **k: self.apply(getattr(Cmd, name)(*a, **k))
        raise AttributeError(name)

    def set_voice(self, idx, group=-1, position=None, velocity=None, gain=None, active=None):
        def fp(x):
            return C.byref(C.c_float(x)) if x is not None else None
        act = C.byref(C.c_int(int(active))) if active is not None else None
        for m in range(self.group.n):                     # every member keeps the whole state
            c = self.group.lib.blast_group_conductor_member(self.h, m)
            check(self.group.lib.blast_conductor_set_voice(c, group, idx, fp(position), fp(velocity), fp(gain), act))

    def get_voice(self, idx, group=-1, member=None):
        """state of a voice as held by the member that renders it (positions advance only there)"""
        out = []
        for m in range(self.group.n) if member is None else [member]:
            s = _lib.VoiceState()
            c = self.group.lib.blast_group_conductor_member(self.h, m)
            check(self.group.lib.blast_conductor_get_voice(c, group, idx, C.byref(s)))
            out.append(s)
        return out

    def coordinate(self, frames: int) -> np.ndarray:
        bus = np.zeros(frames * self.out_channels, dtype=np.int16)
        check(self.group.lib.blast_group_conductor_coordinate(self.h, frames, bus.ctypes.data if bus.size else None))
        return bus
