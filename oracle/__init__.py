"""ctypes binding of the CPU oracle (oracle/blast_oracle.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under audio_decoder_b200/ imports this.
Parity status: "parity unpinned" by the reference's own artefacts (see blast_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libblast_oracle.so")

OK, IO, UNSUPPORTED_FORMAT, UNEXPECTED_EOF, INVALID_DATA, REF_PANIC, BAD_ARG = 0, 1, 2, 3, 4, 5, 101

TM_PROCESS, TM_VOICE, TM_GROUP, TM_CONTEXT, TM_TBD = range(5)
TU_SAMPLES, TU_MILLIS, TU_BPM = range(3)
(CMD_LOAD, CMD_START, CMD_PAUSE, CMD_RESUME, CMD_STOP, CMD_UNLOAD, CMD_VELOCITY, CMD_GROUP, CMD_TC, CMD_SEQ,
 CMD_QUIT) = range(11)
IDX_TEMPO, IDX_VOICE, IDX_PROCESS, IDX_GROUP = range(4)


def build(force: bool = False) -> str:
    """Compile the oracle with g++ (a few seconds).  Building the checker is not using it."""
    src = [os.path.join(_HERE, f) for f in ("blast_oracle.cpp", "blast_oracle.h", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class PcmDesc(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("num_channels", C.c_uint32), ("bits_per_sample", C.c_uint32),
                ("big_endian", C.c_uint32), ("data_off", C.c_uint64), ("data_len", C.c_uint64)]


class X128P(C.Structure):
    _fields_ = [("s0", C.c_uint64), ("s1", C.c_uint64)]


class Track(C.Structure):
    _fields_ = [("samples", C.c_void_p), ("n_samples", C.c_uint64), ("num_channels", C.c_uint32),
                ("sample_rate", C.c_uint32)]


class TempoRepr(C.Structure):
    _fields_ = [("idx", C.c_uint64), ("owned", C.c_uint32), ("mode", C.c_uint32), ("unit", C.c_uint32),
                ("interval", C.c_float)]


class Command(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("idx_kind", C.c_uint32), ("idx", C.c_uint64), ("val", C.c_float),
                ("tempo", TempoRepr),
                ("n_members", C.c_uint32), ("member_voice", C.POINTER(C.c_uint64)),
                ("member_update_tempo", C.POINTER(C.c_uint8)), ("member_n_procs", C.POINTER(C.c_uint32)),
                ("member_proc_ids", C.POINTER(C.c_uint64)),
                ("period", C.c_uint64), ("n_steps", C.c_uint32), ("steps", C.POINTER(C.c_float)),
                ("chance", C.POINTER(C.c_float)), ("rng_s0", C.c_uint64), ("rng_s1", C.c_uint64)]


class VoiceState(C.Structure):
    _fields_ = [("active", C.c_uint32), ("position", C.c_float), ("velocity", C.c_float), ("gain", C.c_float),
                ("end", C.c_uint64), ("channels", C.c_uint32), ("tempo_current", C.c_uint32),
                ("tempo_active", C.c_uint32)]


class MpegHeader(C.Structure):
    _fields_ = [("ok", C.c_uint8), ("err", C.c_uint8), ("version_id", C.c_uint8), ("layer_id", C.c_uint8),
                ("not_protected", C.c_uint8), ("padded", C.c_uint8), ("channel_mode", C.c_uint8),
                ("frame_len_ok", C.c_uint8), ("bitrate", C.c_uint32), ("sr", C.c_double), ("version", C.c_float),
                ("layer", C.c_int32), ("payload_len", C.c_uint64), ("skip", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    u8p, i16p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_int16), C.POINTER(C.c_uint64)
    L.orc_last_error.restype = C.c_char_p
    for name in ("orc_wav_probe", "orc_aiff_probe"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_size_t, C.POINTER(PcmDesc)]
    for name in ("orc_wav_parse", "orc_aiff_parse"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_size_t, C.POINTER(PcmDesc), C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_size_t)]
    L.orc_pcm_decode_fast.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(PcmDesc), C.c_void_p]
    L.orc_pcm_out_len.argtypes = [C.POINTER(PcmDesc)]
    L.orc_pcm_out_len.restype = C.c_size_t
    L.orc_pcm24_unpack.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    L.orc_pcm24_unpack.restype = None
    L.orc_free.argtypes = [C.c_void_p]
    L.orc_free.restype = None
    L.orc_file_name.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
    L.orc_ieee_extended.argtypes = [C.c_void_p]
    L.orc_ieee_extended.restype = C.c_double
    L.orc_f64_as_u32.argtypes = [C.c_double]
    L.orc_f64_as_u32.restype = C.c_uint32
    xp = C.POINTER(X128P)
    L.orc_x128p_new.argtypes = [C.c_uint64, xp]
    L.orc_x128p_new.restype = None
    L.orc_x128p_next_u64.argtypes = [xp]
    L.orc_x128p_next_u64.restype = C.c_uint64
    L.orc_x128p_next_f64.argtypes = [xp]
    L.orc_x128p_next_f64.restype = C.c_double
    L.orc_x128p_next_f32.argtypes = [xp]
    L.orc_x128p_next_f32.restype = C.c_float
    L.orc_x128p_next_i64_range.argtypes = [xp, C.c_int64, C.c_int64]
    L.orc_x128p_next_i64_range.restype = C.c_int64
    L.orc_x128p_fill_u64.argtypes = [xp, C.c_uint64, C.c_void_p]
    L.orc_x128p_fill_u64.restype = None
    L.orc_x128p_fill_range.argtypes = [xp, C.c_int64, C.c_int64, C.c_uint64, C.c_void_p]
    L.orc_x128p_fill_range.restype = None
    L.orc_x128p_discard.argtypes = [xp, C.c_uint64]
    L.orc_x128p_discard.restype = None
    L.orc_x128p_checksum.argtypes = [xp, C.c_int64, C.c_int64, C.c_uint64, u64p, u64p, u64p, u64p]
    L.orc_x128p_checksum.restype = None
    L.orc_convert_interval.argtypes = [C.c_uint32, C.c_uint32, C.c_float]
    L.orc_convert_interval.restype = C.c_float
    L.orc_conductor_new.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(Track), C.c_uint32]
    L.orc_conductor_new.restype = C.c_void_p
    L.orc_conductor_free.argtypes = [C.c_void_p]
    L.orc_conductor_free.restype = None
    L.orc_conductor_apply.argtypes = [C.c_void_p, C.POINTER(Command)]
    L.orc_conductor_coordinate.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    L.orc_conductor_n_voices.argtypes = [C.c_void_p, C.c_int]
    L.orc_conductor_n_groups.argtypes = [C.c_void_p]
    L.orc_conductor_get_voice.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.POINTER(VoiceState)]
    L.orc_conductor_set_voice.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.POINTER(C.c_float),
                                          C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int)]
    L.orc_clock_current.argtypes = [C.c_void_p]
    L.orc_clock_current.restype = C.c_uint64
    L.orc_position_walk.argtypes = [C.c_float, C.c_float, C.c_uint64, C.c_uint64, C.c_void_p]
    L.orc_position_walk.restype = None
    L.orc_mpeg_parse_header.argtypes = [C.c_uint32, C.POINTER(MpegHeader)]
    L.orc_mpeg_parse_header.restype = None
    L.orc_mpeg_match_ref.argtypes = [C.POINTER(MpegHeader), C.POINTER(MpegHeader)]
    L.orc_mpeg_sync_scan.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, u64p]
    L.orc_mpeg_parse.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_uint64, u64p,
                                 C.POINTER(C.c_uint32), u64p, C.c_void_p, C.c_uint64, u64p]
    _lib = L
    return L


class OracleError(Exception):
    def __init__(self, code: int, msg: str):
        super().__init__(f"oracle status {code}: {msg}")
        self.code = code


def _check(rc: int):
    if rc != OK:
        raise OracleError(rc, lib().orc_last_error().decode())


def _bytes_ptr(buf):
    """Return (void* address, length, keepalive) for bytes / numpy uint8."""
    if isinstance(buf, np.ndarray):
        a = np.ascontiguousarray(buf, dtype=np.uint8)
        return a.ctypes.data, a.size, a
    b = bytes(buf)
    a = np.frombuffer(b, dtype=np.uint8)
    return a.ctypes.data, a.size, (a, b)


# ---------------- decode ----------------
def wav_probe(buf) -> PcmDesc:
    p, n, _k = _bytes_ptr(buf)
    d = PcmDesc()
    _check(lib().orc_wav_probe(p, n, C.byref(d)))
    return d


def aiff_probe(buf) -> PcmDesc:
    p, n, _k = _bytes_ptr(buf)
    d = PcmDesc()
    _check(lib().orc_aiff_probe(p, n, C.byref(d)))
    return d


def _parse(fn, buf):
    p, n, _k = _bytes_ptr(buf)
    d = PcmDesc()
    out = C.c_void_p()
    cnt = C.c_size_t()
    _check(fn(p, n, C.byref(d), C.byref(out), C.byref(cnt)))
    if cnt.value:
        arr = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_int16)), shape=(cnt.value,)).copy()
    else:
        arr = np.zeros(0, dtype=np.int16)
    lib().orc_free(out)
    return d, arr


def wav_parse(buf):
    """Faithful wav::parse on an in-memory file image -> (desc, int16 samples)."""
    return _parse(lib().orc_wav_parse, buf)


def aiff_parse(buf):
    return _parse(lib().orc_aiff_parse, buf)


def pcm_decode_fast(buf, desc: PcmDesc) -> np.ndarray:
    p, n, _k = _bytes_ptr(buf)
    out = np.empty(lib().orc_pcm_out_len(C.byref(desc)), dtype=np.int16)
    _check(lib().orc_pcm_decode_fast(p, n, C.byref(desc), out.ctypes.data))
    return out


def pcm24_unpack(payload, big_endian: bool) -> np.ndarray:
    p, n, _k = _bytes_ptr(payload)
    out = np.empty(n // 3, dtype=np.int32)
    lib().orc_pcm24_unpack(p, n // 3, int(big_endian), out.ctypes.data)
    return out


def file_name(path: str) -> str:
    buf = C.create_string_buffer(4096)
    _check(lib().orc_file_name(path.encode(), buf, 4096))
    return buf.value.decode()


def ieee_extended(b10: bytes) -> float:
    a = np.frombuffer(bytes(b10), dtype=np.uint8)
    return lib().orc_ieee_extended(a.ctypes.data)


# ---------------- RNG ----------------
class Rng:
    """X128P (blast_rand.rs:4-60)."""

    def __init__(self, seed: int | None = None, state: tuple[int, int] | None = None):
        self.g = X128P()
        if state is not None:
            self.g.s0, self.g.s1 = state
        else:
            lib().orc_x128p_new(C.c_uint64(seed & (2**64 - 1)), C.byref(self.g))

    @property
    def state(self):
        return (self.g.s0, self.g.s1)

    def next_u64(self):
        return lib().orc_x128p_next_u64(C.byref(self.g))

    def next_f64(self):
        return lib().orc_x128p_next_f64(C.byref(self.g))

    def next_f32(self):
        return lib().orc_x128p_next_f32(C.byref(self.g))

    def next_i64_range(self, lo, hi):
        return lib().orc_x128p_next_i64_range(C.byref(self.g), lo, hi)

    def fill_u64(self, n) -> np.ndarray:
        out = np.empty(n, dtype=np.uint64)
        lib().orc_x128p_fill_u64(C.byref(self.g), n, out.ctypes.data)
        return out

    def fill_range(self, lo, hi, n) -> np.ndarray:
        out = np.empty(n, dtype=np.int64)
        lib().orc_x128p_fill_range(C.byref(self.g), lo, hi, n, out.ctypes.data)
        return out

    def discard(self, n):
        lib().orc_x128p_discard(C.byref(self.g), n)

    def checksum(self, lo, hi, n):
        v = [C.c_uint64() for _ in range(4)]
        lib().orc_x128p_checksum(C.byref(self.g), lo, hi, n, *[C.byref(x) for x in v])
        return tuple(x.value for x in v)


def convert_interval(sample_rate, unit, interval) -> float:
    return lib().orc_convert_interval(sample_rate, unit, interval)


# ---------------- Conductor ----------------
def tempo_repr(idx=0, owned=True, mode=TM_TBD, unit=TU_SAMPLES, interval=0.0) -> TempoRepr:
    return TempoRepr(idx, int(owned), mode, unit, interval)


class Conductor:
    """engine.rs Conductor restated; tracks are int16 numpy arrays (interleaved)."""

    def __init__(self, out_channels: int, sample_rate: int, tracks):
        self._keep = []
        arr = (Track * max(1, len(tracks)))()
        for i, (samples, ch, sr) in enumerate(tracks):
            s = np.ascontiguousarray(samples, dtype=np.int16)
            self._keep.append(s)
            arr[i] = Track(s.ctypes.data, s.size, ch, sr)
        self.out_channels = out_channels
        self.h = lib().orc_conductor_new(out_channels, sample_rate, arr, len(tracks))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().orc_conductor_free(self.h)
                self.h = None
        except Exception:                    # interpreter shutdown
            pass

    def apply(self, cmd: Command):
        _check(lib().orc_conductor_apply(self.h, C.byref(cmd)))

    # convenience builders mirroring the Command variants (commands.rs:86-161)
    def load(self, track_idx, tempo: TempoRepr | None = None):
        c = Command(kind=CMD_LOAD, idx=track_idx, tempo=tempo or tempo_repr())
        self.apply(c)

    def _transport(self, kind, idx_kind, idx):
        self.apply(Command(kind=kind, idx_kind=idx_kind, idx=idx))

    def start(self, idx, idx_kind=IDX_VOICE):
        self._transport(CMD_START, idx_kind, idx)

    def pause(self, idx, idx_kind=IDX_VOICE):
        self._transport(CMD_PAUSE, idx_kind, idx)

    def resume(self, idx, idx_kind=IDX_VOICE):
        self._transport(CMD_RESUME, idx_kind, idx)

    def stop(self, idx, idx_kind=IDX_VOICE):
        self._transport(CMD_STOP, idx_kind, idx)

    def unload(self, idx):
        self.apply(Command(kind=CMD_UNLOAD, idx=idx))

    def velocity(self, idx, val):
        self.apply(Command(kind=CMD_VELOCITY, idx=idx, val=val))

    def tc(self, tempo: TempoRepr):
        self.apply(Command(kind=CMD_TC, tempo=tempo))

    def group(self, tempo: TempoRepr, members):
        """members: list of (voice_idx, update_tempo, [proc ids])"""
        n = len(members)
        mv = (C.c_uint64 * max(1, n))(*[m[0] for m in members])
        mu = (C.c_uint8 * max(1, n))(*[int(m[1]) for m in members])
        mn = (C.c_uint32 * max(1, n))(*[len(m[2]) for m in members])
        flat = [p for m in members for p in m[2]]
        mp = (C.c_uint64 * max(1, len(flat)))(*flat)
        c = Command(kind=CMD_GROUP, tempo=tempo, n_members=n, member_voice=mv, member_update_tempo=mu,
                    member_n_procs=mn, member_proc_ids=mp)
        self.apply(c)

    def seq(self, idx, tempo: TempoRepr, period, steps, chance, rng_state, idx_kind=IDX_VOICE):
        n = len(steps)
        st = (C.c_float * max(1, n))(*steps)
        chn = (C.c_float * max(1, n))(*chance)
        c = Command(kind=CMD_SEQ, idx_kind=idx_kind, idx=idx, tempo=tempo, period=period, n_steps=n, steps=st,
                    chance=chn, rng_s0=rng_state[0], rng_s1=rng_state[1])
        self.apply(c)

    def coordinate(self, frames: int) -> np.ndarray:
        bus = np.zeros(frames * self.out_channels, dtype=np.int16)
        _check(lib().orc_conductor_coordinate(self.h, frames, bus.ctypes.data))
        return bus

    def n_voices(self, group=-1):
        return lib().orc_conductor_n_voices(self.h, group)

    def n_groups(self):
        return lib().orc_conductor_n_groups(self.h)

    def get_voice(self, idx, group=-1) -> VoiceState:
        s = VoiceState()
        _check(lib().orc_conductor_get_voice(self.h, group, idx, C.byref(s)))
        return s

    def set_voice(self, idx, group=-1, position=None, velocity=None, gain=None, active=None):
        def fp(x):
            return C.byref(C.c_float(x)) if x is not None else None
        act = C.byref(C.c_int(int(active))) if active is not None else None
        _check(lib().orc_conductor_set_voice(self.h, group, idx, fp(position), fp(velocity), fp(gain), act))

    def clock(self):
        return lib().orc_clock_current(self.h)


def position_walk(p0, velocity, end, n) -> np.ndarray:
    out = np.empty(n + 1, dtype=np.float32)
    lib().orc_position_walk(p0, velocity, end, n, out.ctypes.data)
    return out


# ---------------- MPEG ----------------
def mpeg_parse_header(h: int) -> MpegHeader:
    o = MpegHeader()
    lib().orc_mpeg_parse_header(h, C.byref(o))
    return o


def mpeg_match_ref(a: MpegHeader, b: MpegHeader) -> bool:
    return bool(lib().orc_mpeg_match_ref(C.byref(a), C.byref(b)))


def mpeg_sync_scan(buf):
    p, n, _k = _bytes_ptr(buf)
    cnt = C.c_uint64()
    _check(lib().orc_mpeg_sync_scan(p, n, None, None, 0, C.byref(cnt)))
    pos = np.empty(cnt.value, dtype=np.uint64)
    hdr = np.empty(cnt.value, dtype=np.uint32)
    _check(lib().orc_mpeg_sync_scan(p, n, pos.ctypes.data, hdr.ctypes.data, cnt.value, C.byref(cnt)))
    return pos, hdr


def mpeg_parse(buf, reference_compat=True, want_payload=True):
    """-> dict(offsets, ref_header, n_candidates, payload)"""
    p, n, _k = _bytes_ptr(buf)
    noff, ncand, plen, ref = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint32()
    _check(lib().orc_mpeg_parse(p, n, int(reference_compat), None, 0, C.byref(noff), C.byref(ref), C.byref(ncand),
                                None, 0, C.byref(plen)))
    offs = np.empty(noff.value, dtype=np.uint64)
    payload = np.empty(plen.value if want_payload else 0, dtype=np.uint8)
    _check(lib().orc_mpeg_parse(p, n, int(reference_compat), offs.ctypes.data, offs.size, C.byref(noff),
                                C.byref(ref), C.byref(ncand), payload.ctypes.data if want_payload else None,
                                payload.size, C.byref(plen)))
    return dict(offsets=offs, ref_header=ref.value, n_candidates=ncand.value, payload=payload)
