#!/usr/bin/env python
"""C3 (4,096 voices x 2^20 frames -> stereo bus) and C4 (2^32 draws / 65,536 streams) device-resident
timings with CUDA events on the launching stream.  Development tool; bench.py is the driver contract."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audio_decoder_b200 as blast  # noqa: E402
from audio_decoder_b200 import _lib, audio_processing as ap, blast_rand as br  # noqa: E402


def c3(ctx, voices, frames, iters, all_unit, gain_one=False):
    L = ctx.lib
    clip_frames = int(np.ceil(1.5 * frames)) + 2
    clip_words = clip_frames * 2
    draws = (clip_words + 3) // 4                       # 4 i16 per u64 draw
    slab = ctx.alloc(voices * draws * 8)
    # per-voice clips = raw X128P output of stream v (stride = draws): generated on the GPU
    s = br.Streams(ctx, voices, draws, seed=0xC30000)
    s.fill_dev(draws, 0, 100, slab.ptr, None, None)
    ctx.sync()
    prm = br.fill(ctx, 0xC3, 0, 1, 2 * voices, 0, 100, ranged=False, checks=False)[0][0]
    tracks, vps = [], []
    for v in range(voices):
        u = float((int(prm[2 * v]) >> 11) * 2.0 ** -53)
        w = float((int(prm[2 * v + 1]) >> 11) * 2.0 ** -53)
        gain = 1.0 if gain_one else float(np.float32(u) * np.float32(2.0 ** -7))
        vel = 1.0 if (v % 2 == 0 or all_unit) else float(np.float32(0.5) + np.float32(w))
        buf = blast.DevBuf.__new__(blast.DevBuf)
        buf.ctx, buf.ptr, buf.nbytes = ctx, slab.ptr + v * draws * 8, draws * 8
        buf.free = lambda: None
        tracks.append(ap.Track(buf, clip_words, 2))
        vps.append(ap.VoiceParams(v, True, 0.0, vel, gain))
    sc = ap.Scene(ctx, tracks, vps, 2)
    part = ctx.alloc(frames * 2 * 4)
    bus = ctx.alloc(frames * 2 * 2)
    times = []
    for it in range(iters + 2):
        sc.set_voices(vps)
        e0, e1 = ctx.event().record(), None
        sc.render_partial_dev(frames, part.ptr)
        ap.finalize_bus(ctx, part.ptr, bus.ptr, frames * 2)
        e1 = ctx.event().record()
        ms = e0.elapsed_ms(e1)
        sc.check()
        if it >= 2:
            times.append(ms)
    ms = float(np.median(times))
    mean_vel = float(np.mean([v.velocity for v in vps]))
    src_bytes = sum(4.0 * frames * v.velocity for v in vps)
    alg = src_bytes + frames * 2 * 2
    host_bus = bus.download(np.int16, frames * 2)
    res = {"workload": f"C3 {voices} voices x {frames} frames, stereo, {'v=1' if all_unit else 'mixed velocities'}" +
                       (", gain 1.0 (the reference's default: integer mix)" if gain_one else ""),
           "ms": round(ms, 4), "gsamples_per_s": round(voices * frames * 2 / ms / 1e6, 1),
           "alg_GB": round(alg / 1e9, 3), "GBps": round(alg / ms / 1e6, 1), "mean_velocity": round(mean_vel, 4),
           "bus_crc": int(np.bitwise_xor.reduce(host_bus.view(np.uint16).astype(np.uint32) * np.arange(1, frames * 2 + 1, dtype=np.uint32)))}
    sc.close()
    return res


def c3_seq(ctx, voices, frames, iters):
    """SURVEY §8 d, C3 second run: the same scene through the Conductor with one Seq process per voice (own Voice tempo,
    a retrigger candidate every `interval` calls, chance 50): K3a event scan + epochs + per-call stepping."""
    clip_frames = int(np.ceil(1.5 * frames)) + 2
    clip_words = clip_frames * 2
    draws = (clip_words + 3) // 4
    slab = ctx.alloc(voices * draws * 8)
    s = br.Streams(ctx, voices, draws, seed=0xC30000)
    s.fill_dev(draws, 0, 100, slab.ptr, None, None)
    ctx.sync()
    prm = br.fill(ctx, 0xC3, 0, 1, 2 * voices, 0, 100, ranged=False, checks=False)[0][0]
    tracks = []
    for v in range(voices):
        buf = blast.DevBuf.__new__(blast.DevBuf)
        buf.ctx, buf.ptr, buf.nbytes = ctx, slab.ptr + v * draws * 8, draws * 8
        buf.free = lambda: None
        tracks.append(ap.Track(buf, clip_words, 2, 48000))
    part = ctx.alloc(frames * 2 * 4)
    times, retrig = [], 0
    for it in range(iters + 1):
        c = ap.Conductor(ctx, 2, 48000, tracks)
        for v in range(voices):
            u = float((int(prm[2 * v]) >> 11) * 2.0 ** -53)
            w = float((int(prm[2 * v + 1]) >> 11) * 2.0 ** -53)
            vel = 1.0 if v % 2 == 0 else float(np.float32(0.5) + np.float32(w))
            # a beat every 24,000..48,000 calls (0.25..0.5 s at 48 kHz stereo), period 4, all four steps armed
            c.load(v, ap.tempo_repr(mode=ap.TM_VOICE, interval=float(24000 + 8 * (v % 3000))))
            c.seq(v, ap.tempo_repr(owned=False, mode=ap.TM_VOICE, idx=v), 4, [0.0, 1.0, 2.0, 3.0], [50.0] * 4,
                  br.seed_state(0xC35E0000 + v))
            c.velocity(v, vel)
            c.start(v)
            c.set_voice(v, gain=float(np.float32(u) * np.float32(2.0 ** -7)))
        c.render_partial_dev(4096, part.ptr)           # warm-up span: the conductor's buffers are allocated here
        ctx.sync()
        e0 = ctx.event().record()
        c.render_partial_dev(frames, part.ptr)
        e1 = ctx.event().record()
        ms = e0.elapsed_ms(e1)
        if it >= 1:
            times.append(ms)
        c.close()
    ms = float(np.median(times))
    return {"workload": f"C3 + one Seq per voice: {voices} voices x {frames} frames through blast_conductor_render_dev",
            "ms": round(ms, 4), "gsamples_per_s": round(voices * frames * 2 / ms / 1e6, 1),
            "note": "includes the host flatten + H2D/D2H of the voice / Seq tables and every chunk's read-back"}


def c4(ctx, n, draws, iters):
    out = {}
    for name, want_raw, want_rng in [("checks_only", False, False), ("raw", True, False), ("raw+ranged", True, True)]:
        s = br.Streams(ctx, n, draws, seed=42)
        raw = ctx.alloc(8 * n * draws) if want_raw else None
        rng = ctx.alloc(8 * n * draws) if want_rng else None
        chk = ctx.alloc(32 * n)
        times = []
        for it in range(iters + 1):
            e0 = ctx.event().record()
            s.fill_dev(draws, 0, 100, raw.ptr if raw else None, rng.ptr if rng else None, chk.ptr)
            e1 = ctx.event().record()
            ms = e0.elapsed_ms(e1)
            if it >= 1:
                times.append(ms)
        ms = float(np.median(times))
        nbytes = 8 * n * draws * (int(want_raw) + int(want_rng))
        out[name] = {"ms": round(ms, 3), "gdraws_per_s": round(n * draws / ms / 1e6, 1), "write_GBps": round(nbytes / ms / 1e6, 1)}
        del raw, rng
    return out


if __name__ == "__main__":
    a = argparse.ArgumentParser()
    a.add_argument("--voices", type=int, default=4096)
    a.add_argument("--frames", type=int, default=1 << 20)
    a.add_argument("--iters", type=int, default=5)
    a.add_argument("--skip-c4", action="store_true")
    a.add_argument("--skip-c3", action="store_true")
    a.add_argument("--only-seq", action="store_true")
    args = a.parse_args()
    with blast.Context(0) as ctx:
        res = {}
        if args.only_seq:
            res["c3_seq"] = c3_seq(ctx, args.voices, args.frames, 2)
        elif not args.skip_c3:
            res["c3_unit"] = c3(ctx, args.voices, args.frames, args.iters, True)
            res["c3_unit_gain1"] = c3(ctx, args.voices, args.frames, args.iters, True, gain_one=True)
            res["c3_mixed"] = c3(ctx, args.voices, args.frames, args.iters, False)
            res["c3_seq"] = c3_seq(ctx, args.voices, args.frames, 2)
        if not args.skip_c4:
            res["c4"] = c4(ctx, 65536, 65536, 3)
        print(json.dumps(res, indent=1))
