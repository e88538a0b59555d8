#!/usr/bin/env python
"""Generates tests/golden/*.npz — small input/output vectors for every part of the hot path.

The reference (Rust) cannot be built or run in the build image and ships no vectors of its own
(SURVEY.md §8 c), so the expected outputs here come from tests/pyref.py: the pure-Python / numpy
restatement written line by line from the Rust sources, independently of oracle/blast_oracle.cpp.
Both the C++ oracle (CPU, tests/test_golden.py) and the CUDA path (GPU, tests/test_golden_gpu.py) must
reproduce these files bit for bit.  Re-run with:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import pyref  # noqa: E402
import synth  # noqa: E402


def decode():
    out = {}
    cases = {
        "wav_even": synth.wav_image(11, 4096),
        "wav_odd_with_tail": np.concatenate([synth.wav_image(12, 1001), np.array([0x5A], np.uint8)]),
        "wav_extensible_91": None,
        "aiff_24bit": synth.aiff_image(13, 3000),
        "aiff_odd_with_tail": np.concatenate([synth.aiff_image(14, 777), np.array([0xA5], np.uint8)]),
    }
    # WAVE_FORMAT_EXTENSIBLE with cb_size > 0: the cursor skips 0+1+...+13 = 91 bytes (wav.rs:124-127)
    import struct
    r15 = np.random.default_rng(15)
    filler = r15.integers(0, 256, size=91, dtype=np.uint8).tobytes()
    pay = r15.integers(0, 256, size=500, dtype=np.uint8).tobytes()
    ext = (b"RIFF" + struct.pack("<I", 0) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 40, 0xFFFE, 2, 48000, 192000, 4, 16) +
           struct.pack("<H", 22) + struct.pack("<HIH", 16, 3, 1) + filler + b"data" + struct.pack("<I", len(pay)) + pay)
    cases["wav_extensible_91"] = np.frombuffer(ext, np.uint8)
    for name, img in cases.items():
        fn = pyref.wav_parse if name.startswith("wav") else pyref.aiff_parse
        d = fn(img.tobytes())
        out[name + "_image"] = img
        out[name + "_samples"] = np.asarray(d["samples"], dtype=np.int16)
        out[name + "_meta"] = np.array([d["sample_rate"], d["num_channels"], d["bits"], d["data_off"], d["data_len"]], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "decode.npz"), **out)


def render():
    rng = np.random.default_rng(21)
    out = {}
    st = rng.integers(-30000, 30000, size=2 * 400, dtype=np.int16)
    mono = rng.integers(-30000, 30000, size=300, dtype=np.int16)
    tri = rng.integers(-30000, 30000, size=3 * 200, dtype=np.int16)
    scenes = {
        # name: (out_channels, frames, [(samples, channels, position, velocity, gain)])
        "stereo_unit": (2, 500, [(st, 2, 0.0, 1.0, 0.8)]),
        "mixed": (2, 350, [(st, 2, 0.0, 1.0, 1.0), (st, 2, 3.25, 0.73, 0.5), (mono, 1, 0.0, 1.0, 1.3), (mono, 1, 10.0, 0.31, 2.5),
                           (tri, 3, 0.0, 1.0, 1.0), (st, 2, 399.0, -1.0, 1.0)]),
        "mono_bus": (1, 300, [(st, 2, 0.0, 1.0, 1.0), (mono, 1, 0.0, 1.5, 0.9)]),
        "quad_bus": (4, 150, [(tri, 3, 0.0, 0.9, 1.0), (mono, 1, 0.0, 1.0, 1.0), (st, 2, 0.5, 1.0, 4.0)]),
        "saturate_wrap": (2, 64, [(np.full(256, 30000, np.int16), 2, 0.0, 1.0, 2.0), (np.full(256, 30000, np.int16), 2, 0.0, 1.0, 1.0),
                                  (np.full(256, 30000, np.int16), 2, 0.0, 1.0, 1.0)]),
    }
    for name, (oc, frames, voices) in scenes.items():
        pv = [pyref.PyVoice(s, c, p, v, g) for s, c, p, v, g in voices]
        bus = pyref.render(pv, oc, frames)
        out[name + "_bus"] = bus
        out[name + "_cfg"] = np.array([oc, frames, len(voices)], dtype=np.int64)
        out[name + "_final_pos"] = np.array([v.position for v in pv], dtype=np.float32)
        for k, (s, c, p, v, g) in enumerate(voices):
            out[f"{name}_v{k}_samples"] = s
            out[f"{name}_v{k}_params"] = np.array([c, p, v, g], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "render.npz"), **out)


def rng():
    out = {}
    for seed in (0, 1, 42, 0xDEADBEEFCAFEBABE):
        g = pyref.X128P(seed)
        out[f"seed{seed:x}_state"] = np.array([g.s0, g.s1], dtype=np.uint64)
        out[f"seed{seed:x}_u64"] = np.array([g.next_u64() for _ in range(64)], dtype=np.uint64)
        g = pyref.X128P(seed)
        out[f"seed{seed:x}_range_0_100"] = np.array([g.next_i64_range(0, 100) for _ in range(64)], dtype=np.int64)
        g = pyref.X128P(seed)
        out[f"seed{seed:x}_range_50_m7"] = np.array([g.next_i64_range(50, -7) for _ in range(64)], dtype=np.int64)
    # jump-ahead: stream s of (seed 42, stride 1000) = the sequential sequence advanced by 1000 s
    g = pyref.X128P(42)
    seq = np.array([g.next_u64() for _ in range(8 * 1000)], dtype=np.uint64)
    out["jump_seed42_stride1000_first16"] = np.stack([seq[s * 1000:s * 1000 + 16] for s in range(8)])
    np.savez_compressed(os.path.join(HERE, "rng.npz"), **out)


def mpeg():
    out = {}
    streams = {
        "frames": synth.mp3_like(31, 60),
        "ff_flood": np.concatenate([np.zeros(3, np.uint8), np.full(5000, 0xFF, np.uint8), np.zeros(8, np.uint8)]),
        "dense": np.concatenate([np.random.default_rng(32).choice(
            np.array([0xFF, 0xFF, 0xE0, 0xFB, 0x00, 0x90, 0xF3], dtype=np.uint8), size=40000), np.zeros(4, np.uint8)]),
    }
    for name, b in streams.items():
        cands = pyref.mpeg_scan(b.tobytes())
        out[name + "_bytes"] = b
        out[name + "_pos"] = np.array([c[0] for c in cands], dtype=np.uint64)
        out[name + "_hdr"] = np.array([c[1] for c in cands], dtype=np.uint32)
    hdrs = [0xFFFB9064, 0xFFFB9264, 0xFFFA9064, 0xFFF3E0C4, 0xFFF2E0C4, 0xFFE3A000, 0xFFE2A000, 0xFFFBF064, 0xFFFB0064,
            0xFFFB9C64, 0xFFFD9064, 0xFFFF9064, 0xFFF99064, 0xFFFB1064, 0xFFFF1004]
    rows = []
    for h in hdrs:
        d = pyref.mpeg_header(h)
        rows.append([h, 0, 0, 0, 0] if d is None else [h, 1, -1 if d["payload"] is None else d["payload"], d["skip"], d["bitrate"]])
    out["header_table"] = np.array(rows, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "mpeg.npz"), **out)


CONDUCTOR_SCRIPT = [
    # (op, args...) — tempo = (idx, owned, mode, unit, interval); modes / units as in blast_time.rs:67-82
    ("tc", (0, True, 3, 0, 6.0)),
    ("load", 0, (0, True, 1, 0, 16.0)),
    ("load", 1, (0, False, 1, 0, 0.0)),                       # borrows voice 0's tempo: both tick it
    ("load", 2, (0, True, 4, 2, 180000.0)),                   # TBD tempo, BPM -> 16 samples
    ("load", 3, (0, True, 1, 1, 0.5)),                        # millis -> 24 samples
    ("seq", 0, (0, False, 1, 0, 0.0), 4, [0.0, 2.0], [100.0, 50.0], 11),
    ("seq", 1, (0, True, 0, 0, 9.0), 3, [1.0], [100.0], 12),  # own Process tempo
    ("seq", 2, (0, False, 3, 0, 0.0), 2, [0.0], [100.0], 13), # context tempo (never ticked)
    ("seq", 3, (3, False, 1, 0, 0.0), 5, [0.0, 1.0, 4.0], [100.0, 0.0, 75.0], 14),
    ("velocity", 1, 0.75), ("velocity", 3, 1.5),
    ("start", 0), ("start", 1), ("start", 2), ("start", 3),
    ("render", 300),
    ("tstart", 0), ("render", 40), ("tstop", 0),
    ("group", (0, True, 2, 0, 10.0), [(2, True, [0]), (0, False, [])]),
    ("gstart", 0), ("render", 200),
    ("pause", 0), ("velocity", 1, -1.0), ("render", 1), ("resume", 0), ("render", 150),
    ("gstop", 0), ("unload", 1), ("render", 64),
]


def conductor():
    rng = np.random.default_rng(41)
    tracks = [(rng.integers(-20000, 20000, size=n * ch).astype(np.int16), ch) for n, ch in ((500, 2), (400, 1), (300, 2), (350, 3))]
    c = pyref.PyConductor(2, 48000, tracks)
    out = []
    for op in CONDUCTOR_SCRIPT:
        k, a = op[0], op[1:]
        if k == "render":
            out.append(c.coordinate(a[0]))
        elif k == "seq":
            c.seq(a[0], a[1], a[2], a[3], a[4], (pyref.X128P(a[5]).s0, pyref.X128P(a[5]).s1))
        elif k in ("tstart", "tstop"):
            getattr(c, k[1:])(a[0], "tempo")
        elif k in ("gstart", "gstop"):
            getattr(c, k[1:])(a[0], "group")
        elif k == "group":
            c.group(a[0], a[1])
        else:
            getattr(c, k)(*a)
    d = {"bus": np.concatenate(out)}
    for i, (s, ch) in enumerate(tracks):
        d[f"track{i}"] = s
        d[f"track{i}_channels"] = np.array([ch])
    np.savez_compressed(os.path.join(HERE, "conductor.npz"), **d)


if __name__ == "__main__":
    conductor()
    decode()
    render()
    rng()
    mpeg()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
