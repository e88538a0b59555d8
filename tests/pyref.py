"""Independent second restatement of the reference hot path in pure Python / numpy.

Written from the Rust sources (not from oracle/blast_oracle.cpp) so that the two
restatements can pin each other; small cases only (pure-Python loops).
Citations are relative to /root/reference/blast/src/.
"""
from __future__ import annotations

import math
import struct

import numpy as np

M64 = (1 << 64) - 1
f32 = np.float32


# ---------------- blast_rand.rs:4-60 ----------------
def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def rotl(x, k):
    return ((x << k) | (x >> (64 - k))) & M64


class X128P:
    def __init__(self, seed=None, state=None):
        if state is not None:
            self.s0, self.s1 = state
        else:
            self.s0 = splitmix64(seed & M64)
            self.s1 = splitmix64((seed + 0x9E3779B97F4A7C15) & M64)

    def next_u64(self):
        r = (self.s0 + self.s1) & M64
        s1 = self.s1 ^ self.s0
        self.s0 = rotl(self.s0, 55) ^ s1 ^ ((s1 << 14) & M64)
        self.s1 = rotl(s1, 36)
        return r

    def next_f64(self):
        return float(self.next_u64() >> 11) * (1.0 / float(1 << 53))

    def next_f32(self):
        return f32(self.next_f64())

    def next_i64_range(self, lower, upper):
        r = self.next_u64()
        rng = upper - lower if upper > lower else lower - upper
        val = (r * rng) >> 64
        out = (lower + val) & M64
        return out - (1 << 64) if out >= (1 << 63) else out


# GF(2) view of the state transition: state = s0 | s1 << 64 (128-bit int)
def _step_state(st: int) -> int:
    g = X128P(state=(st & M64, st >> 64))
    g.next_u64()
    return g.s0 | (g.s1 << 64)


def transition_columns():
    """columns of T: T*e_i for i in 0..127"""
    return [_step_state(1 << i) for i in range(128)]


def mat_vec(cols, v):
    out = 0
    i = 0
    while v:
        if v & 1:
            out ^= cols[i]
        v >>= 1
        i += 1
    return out


def mat_mul(a, b):
    """(a*b) columns: a applied to each column of b"""
    return [mat_vec(a, c) for c in b]


def mat_pow(cols, n):
    result = [1 << i for i in range(128)]
    base = cols
    while n:
        if n & 1:
            result = mat_mul(base, result)
        base = mat_mul(base, base)
        n >>= 1
    return result


def jump_poly(state, poly):
    """canonical polynomial jump: for each set bit of poly (s0 word first) xor in the state, stepping each bit"""
    g = X128P(state=state)
    a0 = a1 = 0
    for word in poly:
        for b in range(64):
            if (word >> b) & 1:
                a0 ^= g.s0
                a1 ^= g.s1
            g.next_u64()
    return a0, a1


# ---------------- Rust casts ----------------
def f32_as_i16(x) -> int:
    x = float(x)
    if math.isnan(x):
        return 0
    if x >= 32767.0:
        return 32767
    if x <= -32768.0:
        return -32768
    return int(x)


def f32_as_usize(x) -> int:
    x = float(x)
    if math.isnan(x) or x <= 0:
        return 0
    if x >= 2.0 ** 64:
        return (1 << 64) - 1
    return int(x)


def wrap_i16(x: int) -> int:
    x &= 0xFFFF
    return x - 0x10000 if x >= 0x8000 else x


# ---------------- engine.rs:46-81, 386-448 (static voices, no processes) ----------------
class PyVoice:
    def __init__(self, samples, channels, position=0.0, velocity=1.0, gain=1.0, active=True):
        self.samples = [int(s) for s in samples]
        self.channels = channels
        self.end = len(self.samples) // channels - 1          # engine.rs:302
        self.position = f32(position)
        self.velocity = f32(velocity)
        self.gain = f32(gain)
        self.active = active

    def process(self, acc: int, ch: int) -> int:
        if not self.active:
            return acc
        idx = f32_as_usize(self.position)
        if idx >= self.end:
            return acc
        C = self.channels
        if C == 1:
            if ch < 2:
                ch = 0
            else:
                return acc
        elif ch >= C:
            return acc
        s0 = f32(self.samples[idx * C + ch % C])
        if self.velocity != f32(1.0):
            frac = f32(self.position - np.trunc(self.position))
            s1 = f32(self.samples[(idx + 1) * C + ch % C])
            sample = f32(f32(s0 * f32(f32(1.0) - frac)) + f32(s1 * frac))
        else:
            sample = s0
        acc = wrap_i16(acc + f32_as_i16(f32(sample * self.gain)))
        if ch == C - 1:
            self.position = f32(self.position + self.velocity)
        return acc


def render(voices, out_channels, frames):
    out = []
    with np.errstate(all="ignore"):
        for _f in range(frames):
            for ch in range(out_channels):
                acc = 0
                for v in voices:
                    acc = v.process(acc, ch)
                out.append(acc)
    return np.array(out, dtype=np.int16)


# ---------------- wav.rs:69-154 / aiff.rs:99-170 ----------------
class Eof(Exception):
    pass


class Unsupported(Exception):
    pass


class Invalid(Exception):
    pass


def _take(b, pos, n):
    if pos + n > len(b):
        raise Eof()
    return b[pos:pos + n], pos + n


def wav_parse(b: bytes):
    pos = 0
    _, pos = _take(b, pos, 4)
    _, pos = _take(b, pos, 4)
    _, pos = _take(b, pos, 4)
    _, pos = _take(b, pos, 4)
    x, pos = _take(b, pos, 4)
    fmt_size = int.from_bytes(x, "little")
    x, pos = _take(b, pos, 2)
    tag = int.from_bytes(x, "little")
    if tag not in (1, 3, 6, 7, 0xFFFE):
        raise Unsupported()
    x, pos = _take(b, pos, 2)
    ch = int.from_bytes(x, "little")
    x, pos = _take(b, pos, 4)
    rate = int.from_bytes(x, "little")
    _, pos = _take(b, pos, 4)
    _, pos = _take(b, pos, 2)
    x, pos = _take(b, pos, 2)
    bits = int.from_bytes(x, "little")
    if fmt_size >= 18:
        x, pos = _take(b, pos, 2)
        if int.from_bytes(x, "little") > 0:
            _, pos = _take(b, pos, 8)
            pos += sum(range(14))
    _, pos = _take(b, pos, 4)
    x, pos = _take(b, pos, 4)
    n = int.from_bytes(x, "little")
    samples = []
    for i in range(pos, pos + n, 2):
        if i + 1 >= len(b):
            raise Eof()
        samples.append(struct.unpack("<h", b[i:i + 2])[0])
    return dict(sample_rate=rate, num_channels=ch, bits=bits, data_off=pos, data_len=n,
                samples=np.array(samples, dtype=np.int16))


def ieee_extended(bs: bytes) -> float:
    sign = bs[0] & 0x80
    exp = ((bs[0] & 0x7F) << 8) | bs[1]
    mant = int.from_bytes(bs[2:10], "big")
    if exp == 0 and mant == 0:
        return 0.0
    if exp == 0x7FFF:
        return (-math.inf if sign else math.inf) if mant == 0 else math.nan
    e = exp - 16383 - 63
    try:
        val = float(mant) * (2.0 ** e)
    except OverflowError:
        val = math.inf
    return -val if sign else val


def f64_as_u32(x: float) -> int:
    if math.isnan(x) or x <= 0:
        return 0
    if x >= 4294967295.0:
        return 0xFFFFFFFF
    return int(x)


def aiff_parse(b: bytes):
    pos = 0
    _, pos = _take(b, pos, 4)
    _, pos = _take(b, pos, 4)
    _, pos = _take(b, pos, 4)
    _, pos = _take(b, pos, 4)
    x, pos = _take(b, pos, 4)
    if int.from_bytes(x, "big") != 18:
        raise Invalid()
    x, pos = _take(b, pos, 2)
    ch = int.from_bytes(x, "big")
    _, pos = _take(b, pos, 4)
    x, pos = _take(b, pos, 2)
    bits = int.from_bytes(x, "big")
    x, pos = _take(b, pos, 10)
    rate = f64_as_u32(ieee_extended(x))
    _, pos = _take(b, pos, 4)
    x, pos = _take(b, pos, 4)
    n = (int.from_bytes(x, "big") - 8) & 0xFFFFFFFF
    _, pos = _take(b, pos, 8)
    samples = []
    for i in range(pos, pos + n, 2):
        if i + 1 >= len(b):
            raise Eof()
        samples.append(struct.unpack(">h", b[i:i + 2])[0])
    return dict(sample_rate=rate, num_channels=ch, bits=bits, data_off=pos, data_len=n,
                samples=np.array(samples, dtype=np.int16))


# ---------------- mpeg.rs ----------------
def mpeg_scan(b: bytes):
    out = []
    cur, n = 0, len(b)
    while cur < n:
        if b[cur] == 0xFF and (b[cur + 1] & 0xE0) == 0xE0:      # IndexError == the reference's panic
            if cur + 3 >= n:
                break
            out.append((cur, int.from_bytes(b[cur:cur + 4], "big")))
            cur += 4
        else:
            cur += 1
    return out


def mpeg_header(h: int):
    """-> None on Err, else dict"""
    b1, b2, b3 = (h >> 16) & 0xFF, (h >> 8) & 0xFF, h & 0xFF
    version = (((b1 >> 4) & 1) << 1) | (b1 & 1)
    if version == 1:
        return None
    layer = (b1 >> 1) & 3
    if layer == 0:
        return None
    not_prot = b1 & 1
    e = b2 >> 4
    if e in (0, 15):
        return None
    bitrate = [8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160][e - 1]
    base = {3: 32000.0, 2: 16000.0, 0: 8000.0}[version]
    ff = (b2 & 0xF) >> 2
    sr = [base * 1.378125, base * 1.5, base, 0.0][ff]
    if sr == 0.0:
        return None
    padded = (b2 >> 1) & 1
    lay = {1: 3, 2: 2, 3: 1}[layer]
    prot = not_prot == 0
    br = bitrate * 1000.0
    fl = 144.0 * br / sr if lay in (2, 3) else (12.0 * br / sr) * 4.0
    payload = None if fl < 20.0 else int(fl) - (20 if prot else 4) + padded
    return dict(version={0: 2.5, 2: 2.0, 3: 1.0}[version], layer=lay, protected=prot, bitrate=bitrate, sr=sr,
                padded=padded, channel_mode=b3 >> 6, payload=payload, skip=6 if prot else 4, fl=fl)


# ---------------- engine.rs:26-248, 318-384, 451-542; blast_time.rs:58-161; processes.rs:52-99 ----------------
# The Command-driven Conductor, written from the Rust sources (not from oracle/blast_oracle.cpp).  Objects and
# method names follow the reference; `Rc<RefCell<TempoState>>` sharing is plain Python object identity.
TM_PROCESS, TM_VOICE, TM_GROUP, TM_CONTEXT, TM_TBD = range(5)
TU_SAMPLES, TU_MILLIS, TU_BPM = range(3)
M32 = (1 << 32) - 1


class RefPanic(Exception):
    """the reference would panic here (unwrap on None / index out of bounds)"""


def f32_as_i64(x) -> int:
    x = float(x)
    if math.isnan(x):
        return 0
    if x >= 9223372036854775807.0:
        return (1 << 63) - 1
    if x <= -9223372036854775808.0:
        return -(1 << 63)
    return int(x)


class PyTempo:                                   # blast_time.rs:58-145
    def __init__(self, sample_rate):
        self.sample_rate = sample_rate
        self.mode, self.unit = TM_TBD, TU_SAMPLES        # TempoState::new(None)
        self.interval = f32(sample_rate)
        self.active = False
        self.current = 0

    def init(self, mode, unit, interval):        # blast_time.rs:99-104 + convert_interval :151-161
        interval = f32(interval)
        with np.errstate(all="ignore"):
            if unit == TU_MILLIS:
                interval = f32(f32(self.sample_rate) * f32(interval / f32(1000.0)))
            elif unit == TU_BPM:
                interval = f32(f32(self.sample_rate) * f32(f32(60.0) / interval))
        self.mode, self.unit, self.interval = mode, unit, interval

    def update(self):                            # update(1.0): current += 1.0 as u32, wrapping in release builds
        self.current = (self.current + 1) & M32

    def cur(self):                               # current(): current as f32 / interval
        with np.errstate(all="ignore"):
            return f32(f32(self.current) / self.interval)

    def reset(self):
        self.current = 0

    def start(self):
        self.reset()
        self.active = True

    def stop(self):
        self.active = False
        self.reset()


class PySeq:                                     # processes.rs:52-99
    def __init__(self, tempo, period, steps, chance, rng):
        self.active, self.tempo, self.period = True, tempo, period
        self.steps, self.chance, self.rng, self.idx = [f32(s) for s in steps], [f32(c) for c in chance], rng, 0

    def process(self, voice):
        if not self.active or not self.tempo.active:
            return
        with np.errstate(all="ignore"):
            current = f32(np.fmod(self.tempo.cur(), f32(self.period)))     # Rust `%` on f32 = fmodf
        if self.idx >= len(self.steps):
            raise RefPanic("steps index out of bounds")
        if current == self.steps[self.idx]:
            rand = self.rng.next_i64_range(0, 100)
            if rand < f32_as_i64(self.chance[self.idx]):
                voice.position = f32(0.0) if voice.velocity >= f32(0.0) else f32(voice.end)
            self.idx = (self.idx + 1) % len(self.steps)


class PyCVoice:                                  # engine.rs:279-448
    def __init__(self, samples, channels, tempo):
        self.samples = [int(s) for s in samples]
        self.channels = channels
        self.end = len(self.samples) // channels - 1
        self.active, self.position, self.velocity, self.gain = False, f32(0.0), f32(1.0), f32(1.0)
        self.tempo, self.processes, self.proc_tempi = tempo, [], []

    def home(self):
        return f32(0.0) if self.velocity >= f32(0.0) else f32(self.end)

    def start(self):
        self.active = True
        for p in self.processes:
            p.idx = 0
        if self.tempo.mode in (TM_VOICE, TM_TBD):
            self.tempo.start()
        for t in self.proc_tempi:
            t.start()
        self.position = self.home()

    def stop(self):
        self.active = False
        for p in self.processes:
            p.idx = 0
        if self.tempo.mode == TM_VOICE:
            self.tempo.stop()
        for t in self.proc_tempi:
            t.active = False
            t.reset()
        self.position = self.home()

    def process(self, acc, ch):
        if not self.active:
            return acc
        for p in self.processes:
            p.process(self)
        if self.tempo.mode in (TM_VOICE, TM_TBD):
            self.tempo.update()
        for t in self.proc_tempi:
            t.update()
        idx = f32_as_usize(self.position)
        if idx >= self.end:
            return acc
        C = self.channels
        if C == 1:
            if ch < 2:
                ch = 0
            else:
                return acc
        elif ch >= C:
            return acc
        with np.errstate(all="ignore"):
            s0 = f32(self.samples[idx * C + ch % C])
            if self.velocity != f32(1.0):
                frac = f32(self.position - np.trunc(self.position))
                s1 = f32(self.samples[(idx + 1) * C + ch % C])
                sample = f32(f32(s0 * f32(f32(1.0) - frac)) + f32(s1 * frac))
            else:
                sample = s0
            acc = wrap_i16(acc + f32_as_i16(f32(sample * self.gain)))
            if ch == C - 1:
                self.position = f32(self.position + self.velocity)
        return acc


class PyGroup:                                   # engine.rs:451-542
    def __init__(self, voices, tempo):
        self.active, self.tempo, self.voices, self.processes = False, tempo, voices, []

    def start(self):
        self.active = True
        if self.tempo.mode == TM_GROUP:
            self.tempo.active = True
            self.tempo.reset()
        for v in self.voices:
            v.start()

    def stop(self):
        self.active = False
        for v in self.voices:
            v.active = False
        if self.tempo.mode == TM_GROUP:
            self.tempo.active = False
            self.tempo.reset()

    def process(self, acc, ch):
        if not self.active:
            return acc
        for v in self.voices:
            acc = v.process(acc, ch)
        if self.tempo.mode == TM_GROUP:
            self.tempo.update()
        return acc


class PyConductor:                               # engine.rs:26-275
    def __init__(self, out_channels, sample_rate, tracks):
        self.out_channels, self.sample_rate = out_channels, sample_rate
        self.tracks = tracks                     # [(samples, channels)]
        self.voices, self.groups, self.tempo_cons = [], [], []

    def tempo_from_repr(self, idx, owned, mode, unit, interval):
        t = PyTempo(self.sample_rate)
        try:
            if owned:
                t.init(mode, unit, interval)
            elif mode == TM_VOICE:
                t = self.voices[idx].tempo
            elif mode == TM_GROUP:
                t = self.groups[idx].tempo
            elif mode == TM_CONTEXT:
                t = self.tempo_cons[idx]
        except IndexError:
            raise RefPanic("tempo index")
        return t

    def load(self, track_idx, tempo):
        if track_idx >= len(self.tracks):
            raise RefPanic("track")
        t = self.tempo_from_repr(*tempo)
        s, ch = self.tracks[track_idx]
        self.voices.append(PyCVoice(s, ch, t))

    def _target(self, idx_kind, idx):
        try:
            return {"voice": self.voices, "group": self.groups, "tempo": self.tempo_cons}[idx_kind][idx]
        except IndexError:
            raise RefPanic(idx_kind)

    def start(self, idx, idx_kind="voice"):
        self._target(idx_kind, idx).start()

    def stop(self, idx, idx_kind="voice"):
        self._target(idx_kind, idx).stop()

    def pause(self, idx, idx_kind="voice"):
        self._target(idx_kind, idx).active = False

    def resume(self, idx, idx_kind="voice"):
        self._target(idx_kind, idx).active = True

    def unload(self, idx):
        if idx >= len(self.voices):
            raise RefPanic("unload")
        self.voices.pop(idx)

    def velocity(self, idx, val):
        if idx >= len(self.voices):
            raise RefPanic("velocity")
        self.voices[idx].velocity = f32(val)

    def group(self, tempo, members):
        t = self.tempo_from_repr(*tempo)
        moved = []
        for idx, update_tempo, p_ids in members:
            if idx >= len(self.voices):
                raise RefPanic("group member")
            v = self.voices.pop(idx)
            if update_tempo:
                v.tempo = t
                for p in p_ids:
                    v.processes[p].tempo = t
            moved.append(v)
        self.groups.append(PyGroup(moved, t))

    def tc(self, tempo):
        self.tempo_cons.append(self.tempo_from_repr(*tempo))

    def seq(self, idx, tempo, period, steps, chance, rng_state, idx_kind="voice"):
        t = self.tempo_from_repr(*tempo)
        s = PySeq(t, period, steps, chance, X128P(state=rng_state))
        if idx_kind == "voice":
            if idx >= len(self.voices):
                raise RefPanic("seq voice")
            self.voices[idx].processes.append(s)
            if tempo[2] == TM_PROCESS:
                self.voices[idx].proc_tempi.append(t)
        elif idx_kind == "group":
            if idx >= len(self.groups):
                raise RefPanic("seq group")
            self.groups[idx].processes.append(s)

    def coordinate(self, frames):
        out = []
        for _f in range(frames):
            for ch in range(self.out_channels):
                acc = 0
                for v in self.voices:
                    if v.active:
                        acc = v.process(acc, ch)
                for g in self.groups:
                    if g.active:
                        acc = g.process(acc, ch)
                out.append(acc)
        return np.array(out, dtype=np.int16)
