"""Deterministic synthetic inputs for the BLAST hot path (SURVEY.md §8 d).

Test / bench infrastructure, numpy only.  Payload bytes are seeded uniform random bytes; the
container headers are the canonical 44-byte WAV and 54-byte AIFF preambles the reference's
parsers expect (wav.rs:69-138, aiff.rs:99-154).
"""
from __future__ import annotations

import struct

import numpy as np

AIFF_RATE_48000 = bytes.fromhex("400ebb80000000000000")
AIFF_RATE_44100 = bytes.fromhex("400eac44000000000000")


def wav_header(data_len: int, channels=2, rate=44100, bits=16, tag=1) -> bytes:
    blk = channels * bits // 8
    return (b"RIFF" + struct.pack("<I", (36 + data_len) & 0xFFFFFFFF) + b"WAVE" + b"fmt " +
            struct.pack("<IHHIIHH", 16, tag, channels, rate, rate * blk, blk, bits) + b"data" +
            struct.pack("<I", data_len))


def aiff_header(data_len: int, channels=2, bits=24, rate_bytes=AIFF_RATE_48000) -> bytes:
    frames = data_len // max(1, channels * ((bits + 7) // 8))
    return (b"FORM" + struct.pack(">I", (46 + data_len) & 0xFFFFFFFF) + b"AIFF" + b"COMM" + struct.pack(">I", 18) +
            struct.pack(">HIH", channels, frames, bits) + rate_bytes + b"SSND" +
            struct.pack(">III", data_len + 8, 0, 0))


def payload(seed: int, n: int) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, 256, size=n, dtype=np.uint8)


def wav_image(seed: int, data_len: int, **kw) -> np.ndarray:
    """C1-style file image: canonical header + data_len random bytes."""
    h = np.frombuffer(wav_header(data_len, **kw), dtype=np.uint8)
    return np.concatenate([h, payload(seed, data_len)])


def aiff_image(seed: int, data_len: int, **kw) -> np.ndarray:
    """C2-style file image: 54-byte preamble + data_len random bytes (24-bit BE stereo by default)."""
    h = np.frombuffer(aiff_header(data_len, **kw), dtype=np.uint8)
    return np.concatenate([h, payload(seed, data_len)])


# BASELINE.json configs (SURVEY.md §8 d)
C1_DATA_LEN = 105_840_000            # 10 min, 16-bit stereo 44.1 kHz
C2_FILES = 1024
C2_DATA_LEN = 2_880_000              # 10 s, 24-bit stereo 48 kHz
C3_VOICES = 4096
C3_FRAMES = 1 << 20
C4_STREAMS = 65536
C4_DRAWS = 65536
C5_BYTES = 16 << 30


def mp3_like(seed: int, n_frames: int, tail: int = 1100, frame_payload: int = 413) -> np.ndarray:
    """C5-style stream: frames of 4-byte header (0xFFFB9064 80 % / 0xFFFB9264 20 %) + random payload
    (413 / 414 bytes), then `tail` zero bytes; last byte != 0xFF (mpeg.rs:20 would panic)."""
    rng = np.random.default_rng(seed)
    padded = rng.random(n_frames) < 0.2
    sizes = 4 + frame_payload + padded.astype(np.int64)
    total = int(sizes.sum())
    out = rng.integers(0, 256, size=total + tail, dtype=np.uint8)
    starts = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    out[starts] = 0xFF
    out[starts + 1] = 0xFB
    out[starts + 2] = np.where(padded, 0x92, 0x90)
    out[starts + 3] = 0x64
    out[total:] = 0
    return out


# ---------------------------------------------------------------- SURVEY §8(d) generators (exact, pure Python ints)
_M64 = (1 << 64) - 1


class X128P:
    """blast_rand.rs:4-48 on Python ints: SplitMix64 seeding, xoroshiro128+ (55, 14, 36), next_f64 / next_f32.
    For the handful of per-voice parameters the configs derive from a generator; bulk sample bytes come from the CUDA
    generator (bench) or the C oracle (tests)."""

    def __init__(self, seed: int):
        def splitmix64(x):
            x = (x + 0x9E3779B97F4A7C15) & _M64
            z = x
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
            return z ^ (z >> 31)
        self.s0 = splitmix64(seed & _M64)
        self.s1 = splitmix64((seed + 0x9E3779B97F4A7C15) & _M64)

    def next_u64(self) -> int:
        r = (self.s0 + self.s1) & _M64
        t = self.s1 ^ self.s0
        self.s0 = (((self.s0 << 55) | (self.s0 >> 9)) & _M64) ^ t ^ ((t << 14) & _M64)
        self.s1 = ((t << 36) | (t >> 28)) & _M64
        return r

    def next_f64(self) -> float:
        return (self.next_u64() >> 11) * (1.0 / (1 << 53))

    def next_f32(self) -> np.float32:
        return np.float32(self.next_f64())


def c3_voice_params(n_voices: int = C3_VOICES):
    """per-voice (velocity, gain) from X128P::new(0xC3) in voice order: gain = next_f32() * 2^-7;
    velocity = 1.0 for even v, 0.5 + next_f32() for odd v (f32 arithmetic)"""
    g = X128P(0xC3)
    out = []
    for v in range(n_voices):
        gain = np.float32(g.next_f32() * np.float32(2.0 ** -7))
        vel = np.float32(1.0) if v % 2 == 0 else np.float32(np.float32(0.5) + g.next_f32())
        out.append((float(vel), float(gain)))
    return out


def c3_clip_frames(v: int, frames: int) -> int:
    """clip length of voice v for an N-frame render: N + 2 frames (even v), ceil(1.5 N) + 2 (odd v)"""
    return frames + 2 if v % 2 == 0 else (3 * frames + 1) // 2 + 2
