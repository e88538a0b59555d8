"""The other BASELINE.json configs, device-timed inside bench.py's one JSON line (`configs`): C1 (one 10-minute WAV), C3
(4,096 voices -> one stereo bus: unit velocity, the SURVEY §8 d mix of velocities, + one Seq per voice), C4 (2^32
X128P draws over 65,536 jump-ahead streams: checksums only / raw / raw + ranged), C5 (16 GiB MPEG sync scan / frame
index) and the true packed 24-bit unpack of C2's payloads.  Inputs follow SURVEY §8(d): sample bytes are successive
next_u64 bytes of the X128P seed the survey names, generated in HBM by the library's own stream kernel; per-voice
parameters come from X128P::new(0xC3) (synth.c3_voice_params).  Every entry carries ms, the algorithmic bytes per launch,
achieved GB/s and the fraction of the measured HBM peak; each is checked by a size-independent property, not by the
oracle (tests/ hold the oracle comparisons).

At N > 1 only C3 runs, strong-scaled (voice v -> rank (v + v // N) mod N), with the bus reduced over peer memory inside the render
kernel; `reduce_ms` is what the exchange adds to the same render without it."""
from __future__ import annotations

import ctypes as C

import numpy as np

import synth


def _median_ms(ctx, fn, iters, warm=1):
    times = []
    for it in range(iters + warm):
        e0 = ctx.event().record()
        fn()
        e1 = ctx.event().record()
        ms = e0.elapsed_ms(e1)
        if it >= warm:
            times.append(ms)
    return float(np.median(times))


def _entry(ms, alg_bytes, peak, units, unit_name, **kw):
    gbs = alg_bytes / ms / 1e6 if ms > 0 else 0.0
    d = {"ms": round(ms, 4), "value": round(units / ms / 1e6, 2), "unit": unit_name, "algorithmic_bytes": int(alg_bytes),
         "GBps": round(gbs, 1), "frac": round(gbs / peak, 4) if alg_bytes else None}
    d.update(kw)
    return d


def _max(dist, local, x):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---------------------------------------------------------------------------------------------- C1
def c1(ctx, peak, quick):
    from audio_decoder_b200 import blast_rand as br, file_parsing as fp
    L = ctx.lib
    data_len = synth.C1_DATA_LEN if not quick else synth.C1_DATA_LEN // 16
    hdr = np.frombuffer(synth.wav_header(data_len), dtype=np.uint8)
    words = data_len // 2
    n_rot = 4                                                   # four file images (423 MB > the 126 MB L2), one launch each
    slot = (len(hdr) + data_len + 255) // 256 * 256
    d_pay = ctx.alloc(data_len + 8)
    br.Streams.from_seeds(ctx, [0xC1]).fill_dev((data_len + 7) // 8, 0, 100, d_pay.ptr, None, None)
    d_in, d_out = ctx.alloc(n_rot * slot), ctx.alloc(n_rot * words * 2)
    h = ctx.pinned(256)
    h.u8[:len(hdr)] = hdr
    for k in range(n_rot):
        L.blast_memcpy_h2d(ctx.h, d_in.ptr + k * slot, h.ptr, len(hdr))
        L.blast_memcpy_d2d(ctx.h, d_in.ptr + k * slot + len(hdr), d_pay.ptr, data_len)
    ctx.sync()
    plans = [fp.PcmPlan(ctx, [(d_in.ptr + k * slot + len(hdr), d_out.ptr + k * words * 2, words, False)]) for k in range(n_rot)]
    for p in plans:
        p.run()
    ctx.sync()
    iters = 40
    e0 = ctx.event().record()
    for i in range(iters):
        plans[i % n_rot].run()
    e1 = ctx.event().record()
    ms = e0.elapsed_ms(e1) / iters
    got = d_out.download(np.int16, 1 << 16, offset=2 * (words - (1 << 16)))        # WAV is little-endian: the payload as it lies
    src = d_pay.download(np.int16, 1 << 16, offset=2 * (words - (1 << 16)))
    ok = bool(np.array_equal(got, src))
    for p in plans:
        p.close()
    for b in (d_pay, d_in, d_out, h):
        b.free()
    return _entry(ms, 4 * words, peak, words, "Gsamples/s", workload=f"C1: one {data_len + 44}-byte 16-bit stereo 44.1 kHz WAV "
                  f"({words} samples, X128P::new(0xC1) bytes), payload at +44; 4 rotating file images, one launch per file",
                  kernel="pcm16_decode_batch", check="tail of the decoded samples == payload bytes: " + ("ok" if ok else "FAILED"))


# ---------------------------------------------------------------------------------------------- C2 true 24-bit
def c2_true24(ctx, peak, quick):
    from audio_decoder_b200 import blast_rand as br, file_parsing as fp
    n_files = synth.C2_FILES if not quick else 64
    data_len = synth.C2_DATA_LEN
    draws = data_len // 8
    d_pay = ctx.alloc(n_files * draws * 8)
    br.Streams.from_seeds(ctx, [0xC20000 + i for i in range(n_files)]).fill_dev(draws, 0, 100, d_pay.ptr, None, None)
    n_s = data_len // 3
    out = {}
    for name, kind, width in (("i32", 0, 4), ("i16_top", 1, 2)):
        d_out = ctx.alloc(n_files * n_s * width)
        from audio_decoder_b200 import _lib
        from audio_decoder_b200.errors import check
        # the job array is marshalled once: the timed region is the C call (job table upload + one launch)
        arr = (_lib.Pcm24Job * n_files)(*[_lib.Pcm24Job(d_pay.ptr + i * data_len, d_out.ptr + i * n_s * width, n_s, 1, kind)
                                          for i in range(n_files)])
        ms = _median_ms(ctx, lambda: check(ctx.lib.blast_pcm24_unpack_dev(ctx.h, arr, n_files)), 5, warm=2)
        # property: the top 16 bits of the sign-extended i32 are the big-endian i16 made of the sample's first two bytes
        raw = d_pay.download(np.uint8, 3 * 4096)
        got = d_out.download(np.int32 if kind == 0 else np.int16, 4096)
        want = raw.reshape(-1, 3)[:, :2].copy().view(">i2").reshape(-1).astype(np.int16)
        ok = bool(np.array_equal((got >> 8).astype(np.int16) if kind == 0 else got, want))
        out[name] = _entry(ms, n_files * n_s * (3 + width), peak, n_files * n_s, "Gsamples/s", kernel="pcm24_unpack_batch",
                           check="ok" if ok else "FAILED")
        d_out.free()
    d_pay.free()
    out["workload"] = f"true packed 24-bit unpack of the {n_files} C2 payloads ({n_files * n_s} samples): 3 B in + 4 B (i32) / 2 B (top-16 i16) out"
    return out


# ---------------------------------------------------------------------------------------------- C3
class C3Scene:
    """this rank's voices of the C3 scene: clips from X128P::new(0xC3_0000 + v) bytes in HBM"""

    def __init__(self, ctx, rank, world, n_voices, frames, all_unit):
        import audio_decoder_b200 as blast
        from audio_decoder_b200 import audio_processing as ap, blast_rand as br
        import bench
        self.ctx = ctx
        params = synth.c3_voice_params(n_voices)
        # voice v -> rank (v + v // world) mod world: round robin, rotated by one per row, so that the odd (interpolated,
        # twice as expensive) voices do not all land on the odd ranks as they would with v mod world
        self.ids = [v for v in range(n_voices) if (v + v // world) % world == rank]
        self.slabs = []
        self.src_bytes = 0.0
        track_of = {}
        for parity in (0, 1):                                   # even voices: N + 2 frames; odd: ceil(1.5 N) + 2
            ids = [v for v in self.ids if v % 2 == parity]
            if not ids:
                continue
            clip_frames = synth.c3_clip_frames(parity, frames)
            draws = (clip_frames * 4 + 7) // 8
            slab = ctx.alloc(len(ids) * draws * 8)              # one slab of [voice][draws] u64 rows per clip length
            br.Streams.from_seeds(ctx, [0xC30000 + v for v in ids]).fill_dev(draws, 0, 100, slab.ptr, None, None)
            self.slabs.append(slab)
            for k, v in enumerate(ids):
                track_of[v] = bench.sub_track(blast, ap, ctx, slab.ptr + k * draws * 8, clip_frames * 2, 2)
        self.tracks, self.voices = [], []
        for v in self.ids:                                      # voices in the survey's order: even / odd alternate
            vel, gain = params[v]
            if all_unit:
                vel = 1.0
            self.tracks.append(track_of[v])
            self.voices.append(ap.VoiceParams(len(self.tracks) - 1, True, 0.0, vel, gain))
            self.src_bytes += 4.0 * frames * vel
        ctx.sync()
        self.scene = ap.Scene(ctx, self.tracks, self.voices, 2)
        self.frames = frames

    def close(self):
        self.scene.close()
        for s in self.slabs:
            s.free()


def c3(ctx, rank, world, local, dist, peak, quick, all_unit, reduce="p2p"):
    from audio_decoder_b200 import audio_processing as ap, distributed as bd
    n_voices = synth.C3_VOICES if not quick else 512
    frames = synth.C3_FRAMES if not quick else 1 << 17
    sc = C3Scene(ctx, rank, world, n_voices, frames, all_unit)
    n_slots = frames * 2
    peer = bd.PeerBus(ctx, n_slots, rank, world, fused=(reduce == "p2p"))
    part = ctx.alloc(4 * n_slots)
    bus2 = ctx.alloc(2 * n_slots)

    def fused():
        sc.scene.restore_dev()
        peer.render_reduce(sc.scene, frames)                    # render + tile exchange (two kernels beside the stream, or inside the render kernel)

    def local_only():                                           # the same render without the exchange: K3 + K4 + K5 on this rank's voices
        sc.scene.restore_dev()
        sc.scene.render_partial_dev(frames, part.ptr)
        ap.finalize_bus(ctx, part.ptr, bus2.ptr, n_slots)

    def timed(fn, iters, join=None):
        for _ in range(2):
            fn()
        if join:
            join()
        ctx.sync()
        if dist is not None:
            dist.barrier(device_ids=[local])
        e0 = ctx.event().record()
        for _ in range(iters):
            fn()
        if join:
            join()                                              # the last exchange belongs to the timed region
        e1 = ctx.event().record()
        return _max(dist, local, e0.elapsed_ms(e1) / iters)

    ms = timed(fused, 5, peer.wait)
    peer.check()
    sc.scene.check()
    bus = None
    if rank == 0:
        bus = np.empty(n_slots, dtype=np.int16)
        ctx.lib.blast_memcpy_d2h(ctx.h, bus.ctypes.data, peer.bus_ptr, bus.nbytes)
        ctx.sync()
    ms_local = timed(local_only, 5)
    check = None
    if world == 1:
        # property: the fused finalize == the three-kernel path, bit for bit
        check = "fused bus == K3+K4+K5 bus: " + ("ok" if np.array_equal(bus, bus2.download(np.int16, n_slots)) else "FAILED")
    else:
        # property (linearity, exact): wrap16(sum over ranks of each rank's own finalized bus) == the reduced bus
        import torch
        mine = torch.from_numpy(bus2.download(np.int16, n_slots).astype(np.int32)).cuda()
        dist.all_reduce(mine, op=dist.ReduceOp.SUM)
        if rank == 0:
            check = "reduced bus == wrap16(sum of the ranks' own buses): " + \
                    ("ok" if np.array_equal(bus, mine.cpu().numpy().astype(np.int16)) else "FAILED")
    src = sc.src_bytes
    if dist is not None:
        import torch
        t = torch.tensor([src], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        src = float(t.item())
    alg = src + n_slots * 2
    out = _entry(ms, alg, peak * world, n_voices * frames * 2, "Gsamples/s (voice x frame x channel)",
                 workload=f"C3: {n_voices} stereo voices x {frames} frames -> one stereo S16 bus, "
                          + ("velocity 1.0" if all_unit else "odd voices at velocity 0.5..1.5 (interpolated)")
                          + ", gains next_f32() * 2^-7 from X128P::new(0xC3), clips X128P::new(0xC3_0000 + v)"
                          + (f"; voice v on rank (v + v // {world}) mod {world}" if world > 1 else ""),
                 kernel="voice_position_scan + voice_render_mix_tma + " + ("bus_finalize" if world == 1 else
                        "peer_publish_tiles + bus_reduce_tiles (tile exchange over peer memory)" if reduce == "p2p2" else
                        "tile publish / peer reduce / S16 wrap inside the render kernel"),
                 ms_without_exchange=round(ms_local, 4), reduce_ms=round(ms - ms_local, 4), check=check,
                 frac_note=f"of {world} x the measured HBM peak" if world > 1 else None)
    peer.close()
    sc.close()
    part.free()
    bus2.free()
    return out


def c3_seq(ctx, peak, quick):
    """SURVEY §8 d, C3 second run: the same scene through the Conductor with one Seq process per voice (own Voice tempo,
    a retrigger candidate every `interval` calls, chance 50): K3a event scan + epochs + per-call stepping"""
    from audio_decoder_b200 import audio_processing as ap, blast_rand as br
    n_voices = synth.C3_VOICES if not quick else 512
    frames = synth.C3_FRAMES if not quick else 1 << 17
    sc = C3Scene(ctx, 0, 1, n_voices, frames, False)
    part = ctx.alloc(frames * 2 * 4)
    times = []
    for it in range(3):
        c = ap.Conductor(ctx, 2, 48000, sc.tracks)
        for v, vp in enumerate(sc.voices):
            # a beat every 24,000..48,000 calls (0.25..0.5 s at 48 kHz stereo), period 4, all four steps armed
            c.load(vp.track, ap.tempo_repr(mode=ap.TM_VOICE, interval=float(24000 + 8 * (v % 3000))))
            c.seq(v, ap.tempo_repr(owned=False, mode=ap.TM_VOICE, idx=v), 4, [0.0, 1.0, 2.0, 3.0], [50.0] * 4,
                  br.seed_state(0xC35E0000 + v))
            c.velocity(v, vp.velocity)
            c.start(v)
            c.set_voice(v, gain=vp.gain)
        c.reserve(frames)                                       # the timed span allocates nothing
        ctx.sync()
        e0 = ctx.event().record()
        c.render_partial_dev(frames, part.ptr)
        e1 = ctx.event().record()
        ms = e0.elapsed_ms(e1)
        if it >= 1:
            times.append(ms)
        c.close()
    ms = float(np.median(times))
    out = _entry(ms, sc.src_bytes + frames * 4, peak, n_voices * frames * 2, "Gsamples/s (voice x frame x channel)",
                 workload=f"C3 + one Seq per voice: {n_voices} voices x {frames} frames through blast_conductor_render_dev "
                          "(Voice tempo, interval 24,000..48,000 calls, period 4, chance 50: ~40 retriggers per voice)",
                 kernel="seq_event_scan + voice_position_scan + voice_render_mix_tma + voice_split_fixup",
                 note="whole call: host flatten, table upload, kernels, state read-back; algorithmic bytes as C3 mixed (upper bound: "
                      "retriggers shorten the source spans)")
    sc.close()
    part.free()
    return out


# ---------------------------------------------------------------------------------------------- C4
def c4(ctx, peak, quick):
    from audio_decoder_b200 import blast_rand as br
    n, draws = (synth.C4_STREAMS, synth.C4_DRAWS) if not quick else (8192, 16384)
    out = {"workload": f"C4: {n * draws} draws, {n} streams = X128P::new(42) advanced by s x {draws} (GF(2) jump-ahead)"}
    chk_all = {}
    for name, want_raw, want_rng in (("checks", False, False), ("raw", True, False), ("ranged", True, True)):
        raw = ctx.alloc(8 * n * draws) if want_raw else None
        rng = ctx.alloc(8 * n * draws) if want_rng else None
        chk = ctx.alloc(32 * n)
        state = {}

        def run():
            state["s"] = br.Streams(ctx, n, draws, seed=42)      # fresh states every time: the same draws are generated
            state["s"].fill_dev(draws, 0, 100, raw.ptr if raw else None, rng.ptr if rng else None, chk.ptr)
        run()
        ctx.sync()
        s = br.Streams(ctx, n, draws, seed=42)
        # timed: exactly what the variant names — the checksums (and with them the Lemire product nobody asked for) are
        # only computed where they are the output; the verification run below computes them next to the rows
        t_chk = chk.ptr if not want_raw else None
        ms = _median_ms(ctx, lambda: s.fill_dev(draws, 0, 100, raw.ptr if raw else None, rng.ptr if rng else None, t_chk), 3)
        run()
        chk_all[name] = chk.download(np.uint64, 4 * n)
        nbytes = 8 * n * draws * (int(want_raw) + int(want_rng))
        e = _entry(ms, nbytes, peak, n * draws, "Gdraws/s", kernel="x128p_streams",
                   bound="integer ALU (nothing is written)" if not nbytes else "hbm (write)")
        if want_raw:                                            # property: the materialised rows reproduce the checksums
            row = raw.download(np.uint64, draws, offset=8 * draws * (n - 1))
            ok = int(np.bitwise_xor.reduce(row)) == int(chk_all[name][4 * (n - 1)]) and \
                int(row.sum(dtype=np.uint64)) == int(chk_all[name][4 * (n - 1) + 1])
            e["check"] = "last stream's xor / sum == its checksums: " + ("ok" if ok else "FAILED")
        out["c4_" + name] = e
        for b in (raw, rng, chk):
            if b is not None:
                b.free()
    same = bool(np.array_equal(chk_all["checks"], chk_all["raw"]) and np.array_equal(chk_all["checks"], chk_all["ranged"]))
    out["check"] = "per-stream checksums identical with and without materialising: " + ("ok" if same else "FAILED")
    return out


# ---------------------------------------------------------------------------------------------- C5
def c5(ctx, peak, quick):
    L = ctx.lib
    gib = 16 if not quick else 2
    block_b = 1 << 30
    block = synth.mp3_like(0xC5, block_b // 418 - 4)             # ~1 GiB of frames, zero tail
    block = np.concatenate([block, np.zeros(block_b - block.size, np.uint8)])
    n = gib * block_b
    d = ctx.alloc(n)
    h = ctx.pinned(block_b)
    h.u8[:] = block
    for k in range(gib):
        L.blast_memcpy_h2d(ctx.h, d.ptr + k * block_b, h.ptr, block_b)
    ctx.sync()
    h.free()
    cap = n // 256
    d_pos, d_hdr = ctx.alloc(8 * cap), ctx.alloc(4 * cap)
    cnt = C.c_uint64()

    def scan():
        assert L.blast_mpeg_scan_dev(ctx.h, d.ptr, n, d_pos.ptr, d_hdr.ptr, cap, C.byref(cnt)) == 0, L.blast_last_error()
    ms = _median_ms(ctx, scan, 3)
    out = {"workload": f"C5: {gib} GiB stream (a 1 GiB block of frames 0xFFFB9064 80 % / 0xFFFB9264 20 % + 413/414 random payload bytes, "
                       f"repeated {gib}x), device-resident"}
    # property: the stream is periodic, so are its candidates (the scan is idle across the zero tail of every block)
    per = cnt.value // gib
    p0 = d_pos.download(np.uint64, min(per, 1 << 18))
    pk = d_pos.download(np.uint64, min(per, 1 << 18), offset=8 * per * (gib - 1))
    ok = cnt.value % gib == 0 and bool(np.array_equal(p0 + np.uint64((gib - 1) * block_b), pk))
    out["c5_scan"] = _entry(ms, n + 12 * cnt.value, peak, n, "GB/s scanned", kernel="mpeg_walk + span scan + mpeg_compact",
                            candidates=cnt.value, check="candidates of the last block == those of the first + 15 GiB: " + ("ok" if ok else "FAILED"))
    d_pos.free()
    d_hdr.free()
    noff, ncand, ref = C.c_uint64(), C.c_uint64(), C.c_uint32()
    d_off = ctx.alloc(8 * cap)

    def index():
        assert L.blast_mpeg_index_dev(ctx.h, d.ptr, n, 1, d_off.ptr, cap, C.byref(noff), C.byref(ref), C.byref(ncand)) == 0, L.blast_last_error()
    ms = _median_ms(ctx, index, 3)
    offs = d_off.download(np.uint64, min(noff.value, 1 << 20), offset=8 * max(0, noff.value - (1 << 20)))
    ok = bool(np.all(np.diff(offs.astype(np.int64)) >= 0)) and ref.value in (0xFFFB9064, 0xFFFB9264) and int(offs[-1]) > n - block_b
    out["c5_index"] = _entry(ms, n + 12 * ncand.value + 8 * noff.value, peak, n, "GB/s scanned",
                             kernel="scan + mpeg_hist + mpeg_pick_ref + mpeg_first_pos + mpeg_classify", offsets=noff.value,
                             ref_header=hex(ref.value), check="sorted, reference header, last offset in the last block: " + ("ok" if ok else "FAILED"))
    for b in (d, d_off):
        b.free()
    ctx.trim()
    return out


def run(ctx, rank, world, local, dist, peak, quick=False, reduce="p2p"):
    out = {}
    if world == 1:
        out["c1"] = c1(ctx, peak, quick)
        out["c2_true24"] = c2_true24(ctx, peak, quick)
    out["c3_unit"] = c3(ctx, rank, world, local, dist, peak, quick, True, reduce)
    out["c3_mixed"] = c3(ctx, rank, world, local, dist, peak, quick, False, reduce)
    if world == 1:
        out["c3_seq"] = c3_seq(ctx, peak, quick)
        ctx.trim()
        out.update({k: v for k, v in c4(ctx, peak, quick).items() if k.startswith("c4_")})
        out.update({k: v for k, v in c5(ctx, peak, quick).items() if k.startswith("c5_")})
    return out
