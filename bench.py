#!/usr/bin/env python
"""bench.py — BLAST hot path on B200: PCM decode+mix Gsamples/s (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, libblast_cuda.so)
  python bench.py --impl reference --gpus N ...            # CPU restatement of the reference, all host threads

Workload (config.workload): BASELINE config[1] — batch decode of 1,024 synthetic 24-bit big-endian 48 kHz stereo AIFF
files (2,880,000 payload bytes each = successive next_u64 bytes of X128P::new(0xC2_0000 + file_index), SURVEY §8 d;
reference-exact byte-pair decode into i16 words), followed by the mix of the decoded tracks (stereo voices, velocity 1,
per-voice gain) into one stereo S16 bus.  One "step" = one pass over the whole batch.

Scaling is STRONG: at N > 1 the same 1,024 files are sharded, file i -> rank i mod N (north_star: "1,024 ... files
sharded across GPUs"); the decoded tracks are mixed where they were decoded and the int32 partial buses are reduced
tile by tile inside the render kernel over peer memory (blast_peer_bus).  `weak` is the extra key: every rank decodes
and mixes its own 1,024 files.  Before anything is timed the bus of one step is checked (`bus_check`): the fused
peer-memory path against the unfused render + NCCL all-reduce + finalize, byte for byte, and a 4,096-frame window
against the CPU oracle's Conductor::coordinate over the voices of all ranks.

The JSON line follows the driver contract; `value` is device-resident throughput (CUDA events on the launching
stream), `e2e` the same metric through the host-buffer C-ABI call (pinned host file images in, S16 bus out, copies
inside the timed region; decoded tracks stay in HBM), `e2e_parse_dropin` the variant that also returns every
AudioFile.samples Vec to the host like a literal aiff::parse() drop-in.  `configs` carries the other BASELINE configs
(C1, C3 unit / mixed / +Seq, C4, C5, true 24-bit unpack), each device-timed with its algorithmic bytes and roofline
fraction.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import synth  # noqa: E402

METRIC = "pcm_decode_mix_gsamples_per_s"
UNIT = "Gsamples/s"
CHECK_FRAMES = 4096


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return rank, world, local, dist
    return rank, world, local, None


def max_over_ranks(dist, local, x: float) -> float:
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(dist, local):
    if dist is not None:
        import torch
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize(local)


def c2_gains(n_files: int):
    """per-voice gain of the mix half: f32(next_f32() * 2^-5) from X128P::new(0xC2), in file order (both arms)"""
    g = synth.X128P(0xC2)
    return [float(np.float32(g.next_f32() * np.float32(2.0 ** -5))) for _ in range(n_files)]


def sub_track(blast, ap, ctx, ptr, n_samples, channels, rate=48000):
    buf = blast.DevBuf.__new__(blast.DevBuf)
    buf.ctx, buf.ptr, buf.nbytes = ctx, ptr, n_samples * 2
    buf.free = lambda: None
    return ap.Track(buf, n_samples, channels, rate)


# ------------------------------------------------------------------ our arm: the C2 decode + mix workload
class C2Shard:
    """the files `ids` of the C2 batch on this rank: file images in HBM (256-byte aligned slots, payload at +54) and in
    pinned host memory (back to back), the decode plan, the scene of the decoded tracks"""

    def __init__(self, ctx, ids, data_len, layout, gains, host_images=True):
        import audio_decoder_b200 as blast
        from audio_decoder_b200 import audio_processing as ap, blast_rand as br, file_parsing as fp
        L = ctx.lib
        self.ctx, self.ids, self.data_len = ctx, list(ids), data_len
        n = self.n = len(self.ids)
        hdr = np.frombuffer(synth.aiff_header(data_len), dtype=np.uint8)
        self.image_len = len(hdr) + data_len
        self.slot = slot = (self.image_len + 255) // 256 * 256
        self.words = (data_len + 1) // 2
        self.frames = self.words // 2                                  # stereo
        draws = (data_len + 7) // 8
        # payload bytes = successive next_u64 bytes of X128P::new(0xC2_0000 + file_index), generated by the library's
        # own stream kernel (K6) straight into HBM: [file][draws] rows of u64
        d_pay = ctx.alloc(max(1, n) * draws * 8)
        if n:
            br.Streams.from_seeds(ctx, [0xC20000 + i for i in self.ids]).fill_dev(draws, 0, 100, d_pay.ptr, None, None)
        self.d_in = ctx.alloc(max(1, n) * slot)
        # host images lie BACK TO BACK, like an asset directory read into one buffer: blast_pcm_decode_batch then moves
        # several files per copy (the bytes between two payloads are the next file's own header)
        self.h_in = ctx.pinned(max(1, n) * self.image_len) if host_images else None
        h_hdr = ctx.pinned(256)
        h_hdr.u8[:len(hdr)] = hdr
        for k in range(n):
            L.blast_memcpy_h2d(ctx.h, self.d_in.ptr + k * slot, h_hdr.ptr, len(hdr))
            L.blast_memcpy_d2d(ctx.h, self.d_in.ptr + k * slot + len(hdr), d_pay.ptr + k * draws * 8, data_len)
        if host_images:
            for k in range(n):
                L.blast_memcpy_d2h(ctx.h, self.h_in.ptr + k * self.image_len, self.d_in.ptr + k * slot, self.image_len)
        ctx.sync()
        h_hdr.free()
        self.view = self.h_in.u8.reshape(max(1, n), self.image_len)[:n] if host_images else None
        self.desc = fp.probe("aiff", self.view[0] if host_images and n else np.concatenate([hdr, np.zeros(data_len, np.uint8)]))
        off = self.off = self.desc.data_off
        self.d_out = ctx.alloc(max(1, n) * self.words * 2)
        if layout == "payload":
            self.d_pay = d_pay
            src = [d_pay.ptr + k * draws * 8 for k in range(n)]
        else:
            d_pay.free()
            src = [self.d_in.ptr + k * slot + off for k in range(n)]
        self.plan = fp.PcmPlan(ctx, [(src[k], self.d_out.ptr + k * self.words * 2, self.words, True) for k in range(n)]) if n else None
        self.samples = n * self.words
        self.tracks = [sub_track(blast, ap, ctx, self.d_out.ptr + k * self.words * 2, self.words, 2) for k in range(n)]
        self.voices = [ap.VoiceParams(k, True, 0.0, 1.0, gains[i]) for k, i in enumerate(self.ids)]
        self.scene = ap.Scene(ctx, self.tracks, self.voices, 2)
        self.alg_decode = 4 * self.samples                               # 2 B read + 2 B written per i16 word
        self.alg_mix = 4 * (self.frames - 1) * n                         # source frames touched (+ the S16 bus, once)

    def decode(self):
        if self.plan:
            self.plan.run()

    def close(self):
        if self.plan:
            self.plan.close()
        self.scene.close()
        for b in (self.d_in, self.d_out, self.h_in):
            if b is not None:
                b.free()


def run_ours(args):
    import audio_decoder_b200 as blast
    from audio_decoder_b200 import _lib, audio_processing as ap, distributed as bd

    rank, world, local, dist = dist_setup()
    torch = None
    if world > 1:
        import torch
        ctx = blast.Context(local, stream=torch.cuda.current_stream().cuda_stream)
    else:
        ctx = blast.Context(local)
    L = ctx.lib
    n_files, data_len = args.files, args.data_len
    gains = c2_gains(n_files)
    mix = not args.no_mix
    shard = C2Shard(ctx, range(rank, n_files, world), data_len, args.layout, gains)
    frames = shard.frames
    n_slots = frames * 2
    peak, peak_src = measured_peaks()

    # ---- the mix's exchange step: partial buses reduced tile by tile inside the render kernel over peer memory
    peer = t_part = d_bus2 = None
    if mix:
        if args.reduce in ("p2p", "p2p2"):
            try:
                peer = bd.PeerBus(ctx, n_slots, rank, world, fused=(args.reduce == "p2p"))
            except RuntimeError as e:                                  # raised on EVERY rank or on none
                if rank == 0:
                    print(f"bench.py: {e}; using the NCCL all-reduce instead", file=sys.stderr)
                args.reduce = "nccl"
        if world > 1:
            t_part = torch.empty(n_slots, dtype=torch.int32, device=f"cuda:{local}")
            part2 = t_part.data_ptr()
        else:
            d_part2 = ctx.alloc(4 * n_slots)
            part2 = d_part2.ptr
        d_bus2 = ctx.alloc(2 * n_slots)

    def mix_unfused(sh):
        """the baseline the fused path is measured against: render, NCCL all-reduce of the int32 bus, finalize"""
        sh.scene.restore_dev()
        sh.scene.render_partial_dev(frames, part2)
        if world > 1:
            dist.all_reduce(t_part, op=dist.ReduceOp.SUM)
        ap.finalize_bus(ctx, part2, d_bus2.ptr, n_slots)

    def mix_step(sh):
        if peer is None:
            return mix_unfused(sh)
        sh.scene.restore_dev()
        # K3, K4, then the tile exchange: two kernels on the peer bus's own stream (p2p2) — they overlap the next step's
        # decode; whoever consumes the bus joins them with peer.wait() — or inside K4 (p2p)
        peer.render_reduce(sh.scene, frames)

    bus_ptr = (peer.bus_ptr if peer is not None else d_bus2.ptr) if mix else None

    # ---- bus_check (not timed): fused peer-memory bus == NCCL bus, and a window of it == the CPU oracle
    bus_check = None
    if mix:
        shard.decode()
        mix_step(shard)
        if peer is not None:
            peer.wait()
        a = np.empty(n_slots, dtype=np.int16)
        if rank == 0:
            L.blast_memcpy_d2h(ctx.h, a.ctypes.data, bus_ptr, a.nbytes)
        ctx.sync()
        if peer is not None:
            peer.check()
            mix_unfused(shard)
            b = d_bus2.download(np.int16, n_slots)
            same = bool(np.array_equal(a, b)) if rank == 0 else True
        else:
            same = True
        w = min(CHECK_FRAMES, frames - 2)
        heads = (np.stack([ctx_download(ctx, t.buf.ptr, (w + 2) * 2) for t in shard.tracks]) if shard.n
                 else np.zeros((0, (w + 2) * 2), np.int16))                 # the first w + 2 decoded frames of every track
        if dist is not None:
            box = [None] * world if rank == 0 else None
            dist.gather_object((shard.ids, heads), box, dst=0)
        else:
            box = [(shard.ids, heads)]
        if rank == 0:
            import oracle                                               # the checker, outside every timed region
            order = {}
            for ids, hh in box:
                for i, h in zip(ids, hh):
                    order[i] = h
            oc = oracle.Conductor(2, 48000, [(order[i], 2, 48000) for i in range(n_files)])
            for i in range(n_files):
                oc.load(i)
                oc.set_voice(i, gain=gains[i], active=True)
            exp = oc.coordinate(w)
            window_ok = bool(np.array_equal(a[:2 * w], exp))
            nonzero = int(np.count_nonzero(a))
            bus_check = "ok" if (same and window_ok and nonzero > n_slots // 2) else \
                f"FAILED (fused == nccl: {same}, {w}-frame window == oracle: {window_ok}, nonzero slots {nonzero})"
        if dist is not None:
            box = [bus_check]
            dist.broadcast_object_list(box, src=0)
            bus_check = box[0]
        if bus_check != "ok":
            raise SystemExit(f"bench.py: bus_check {bus_check}")

    def timed(sh, steps, warmup):
        """W warm-up + K timed steps of decode + mix on shard `sh`; CUDA events on the launching stream"""
        n_ev = 3 if mix else 2

        def step(evs=None):
            if evs:
                evs[0].record()
            sh.decode()
            if evs:
                evs[1].record()
            if mix:
                mix_step(sh)
                if evs:
                    evs[2].record()

        for _ in range(warmup):
            step()
        ctx.sync()
        if mix:
            sh.scene.check()
            if peer is not None:
                peer.check()
        barrier(dist, local)
        launches0 = ctx.launch_count
        evs = [[ctx.event() for _ in range(n_ev)] for _ in range(steps)]
        e_end = ctx.event()
        for k in range(steps):
            step(evs[k])
        if mix and peer is not None:
            peer.wait()                                               # the last step's exchange belongs to the timed region
        e_end.record()
        ms_total = evs[0][0].elapsed_ms(e_end)
        ctx.sync()
        barrier(dist, local)
        launches = ctx.launch_count - launches0
        ms_decode = sum(e[0].elapsed_ms(e[1]) for e in evs) / steps
        # (with the exchange on its own stream the mix event closes when the render is enqueued-complete: the exchange of
        # step s is then part of step s + 1's interval, and of the total through the final wait)
        ms_mix = sum(e[1].elapsed_ms(e[2]) for e in evs) / steps if mix else 0.0
        ms_step = max_over_ranks(dist, local, ms_total) / steps
        return dict(ms_step=ms_step, ms_decode=ms_decode, ms_mix=ms_mix, launches=launches,
                    ms_decode_max=max_over_ranks(dist, local, ms_decode), ms_mix_max=max_over_ranks(dist, local, ms_mix))

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t = timed(shard, args.steps, args.warmup)
    ms_step = t["ms_step"]
    total_samples = n_files * shard.words
    value = total_samples / (ms_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (decode moves 2x the bytes of the mix), per-launch average on this rank
    traffic = {}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            traffic = json.load(open(tr))
        except Exception:
            traffic = {}
    achieved = shard.alg_decode / (t["ms_decode"] * 1e-3) / 1e9
    roofline = {"kernel": "pcm16_decode_batch", "bound": "hbm", "achieved": round(achieved, 1), "peak": peak,
                "unit": "GB/s", "frac": round(achieved / peak, 4), "frac_of_nominal_8000": round(achieved / 8000.0, 4),
                "traffic": traffic.get("pcm16_decode_batch") if world == 1 else None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": shard.alg_decode, "ms_per_launch": round(t["ms_decode"], 4),
                "note": "rank 0's shard" if world > 1 else None}
    roofline_mix = None
    alg_mix = shard.alg_mix + 2 * n_slots // world
    if mix:
        ach = alg_mix / (t["ms_mix"] * 1e-3) / 1e9
        roofline_mix = {"kernel": ("voice_position_scan + voice_render_mix_tma + bus_finalize" if world == 1 else
                                   "voice_position_scan + voice_render_mix_tma (render, tile publish, peer reduce, S16 wrap in one kernel)"
                                   if args.reduce == "p2p" else "voice_position_scan + voice_render_mix_tma + peer_publish_tiles + bus_reduce_tiles (tile exchange over peer memory)")
                        if peer is not None else "voice_position_scan + voice_render_mix_tma + NCCL all-reduce(int32) + bus_finalize",
                        "bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                        "algorithmic_bytes_per_step": alg_mix, "ms_per_step": round(t["ms_mix"], 4),
                        "traffic": traffic.get("voice_render_mix_tma_c2") if world == 1 else None}
    alg_step = shard.alg_decode + (alg_mix if mix else 0)
    roofline_step = {"bound": "hbm", "achieved": round(alg_step / (ms_step * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(alg_step / (ms_step * 1e-3) / 1e9 / peak, 4), "algorithmic_bytes_per_step": alg_step,
                     "note": "decode + mix of one rank, max-over-ranks step time"}

    # ---- weak scaling as the extra key (N > 1): every rank decodes and mixes its own 1,024 files
    weak = None
    if world > 1 and not args.no_weak:
        full = C2Shard(ctx, range(n_files), data_len, args.layout, gains, host_images=False)
        tw = timed(full, max(3, args.steps // 2), 3)
        weak = {"value": round(world * full.samples / (tw["ms_step"] * 1e-3) / 1e9, 3), "unit": UNIT,
                "ms_per_step": round(tw["ms_step"], 4), "files_per_gpu": n_files,
                "kernel_ms": {"decode": round(tw["ms_decode_max"], 4), "mix": round(tw["ms_mix_max"], 4)}}
        full.close()

    # ---- e2e: the same step through the host-buffer C ABI.  Pinned host file images -> blast_pcm_decode_batch
    #      (H2D inside) -> render of the decoded tracks -> S16 bus copied back to the host.
    #      "e2e":        AudioFile.samples stay in HBM (their only consumer is the render kernel; SURVEY §8 b)
    #      "e2e_parse_dropin": additionally every AudioFile.samples Vec is delivered to the host, as a literal
    #                    aiff::parse() drop-in must (doubles the PCIe traffic)
    e2e = e2e_dropin = None
    if not args.no_e2e:
        n = shard.n
        wpf = shard.words
        h_out = ctx.pinned(max(1, n) * wpf * 2)
        h_bus = ctx.pinned(2 * n_slots) if mix else None
        files = (C.c_void_p * max(1, n))(*[shard.h_in.ptr + k * shard.image_len for k in range(n)])
        lens = (C.c_size_t * max(1, n))(*([shard.image_len] * n))
        dd = (_lib.PcmDesc * max(1, n))(*([shard.desc] * n))
        host_out = (C.c_void_p * max(1, n))(*[h_out.ptr + k * wpf * 2 for k in range(n)])
        dev_out = (C.c_void_p * max(1, n))(*[shard.d_out.ptr + k * wpf * 2 for k in range(n)])
        pcie = {}
        pp = os.path.join(ROOT, "profiles", "r02_pcie_probe.json")
        if os.path.exists(pp):
            pcie = json.load(open(pp))

        def e2e_step(to_host):
            rc = L.blast_pcm_decode_batch(ctx.h, n, files, lens, dd, host_out if to_host else None, dev_out)
            if rc != 0:
                raise RuntimeError(L.blast_last_error().decode())
            if mix:
                mix_step(shard)
                if peer is not None:
                    peer.wait()
                if rank == 0:                                          # the bus exists on the root
                    L.blast_memcpy_d2h(ctx.h, h_bus.ptr, bus_ptr, 2 * n_slots)
                ctx.sync()

        def e2e_run(to_host):
            for _ in range(2):
                e2e_step(to_host)
            barrier(dist, local)
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                e2e_step(to_host)
            ctx.sync()
            dt = (time.perf_counter() - t0) / args.e2e_steps
            dt = max_over_ranks(dist, local, dt)
            h2d = n * data_len
            d2h = (n * wpf * 2 if to_host else 0) + (2 * n_slots if mix and rank == 0 else 0)
            r = {"value": round(total_samples / dt / 1e9, 3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                 "d2h_bytes_per_step": d2h, "ms_per_step": round(dt * 1e3, 3),
                 "pcie_GBps_per_gpu": round(max(h2d, d2h) / dt / 1e9, 2),
                 "bytes_note": "per rank (rank 0)" if world > 1 else None}
            if pcie:
                r["pcie_peak_GBps"] = pcie.get("bidir_each_GBps" if to_host else "h2d_GBps")
                r["pcie_frac"] = round(r["pcie_GBps_per_gpu"] / r["pcie_peak_GBps"], 3)
                r["pcie_peak_source"] = "tools/pcie_probe.py on this pool's B200 box (profiles/r02_pcie_probe.json)"
            return r

        e2e = e2e_run(False)
        e2e["api"] = ("blast_pcm_decode_batch (pinned host file images in, decoded tracks kept in HBM)" +
                      (" + blast_scene_render_reduce_dev + S16 bus D2H" if mix else ""))
        e2e_dropin = e2e_run(True)
        e2e_dropin["api"] = e2e["api"].replace("decoded tracks kept in HBM", "host AudioFile.samples out AND tracks kept in HBM")
        if n:   # spot-check the e2e result against numpy (not timed)
            got = h_out.view(np.int16, wpf, 0)
            assert np.array_equal(got, shard.view[0, shard.off:shard.off + data_len].view(">i2").astype(np.int16)), "e2e output mismatch"
        h_out.free()
    clocks = sampler.stop() if rank == 0 else None

    workload = (f"C2: batch decode {n_files} x 24-bit BE 48 kHz stereo AIFF ({data_len} payload B each, X128P::new(0xC2_0000 + i) bytes), "
                "reference-exact byte-pair decode to i16" + (f", then mix of the {n_files} decoded tracks (stereo voices, velocity 1, "
                                                             "per-voice gain) into one stereo S16 bus" if mix else ""))
    out = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "i16", "data": "synthetic",
        "config": {"workload": workload, "files_total": n_files, "files_per_gpu": shard.n,
                   "samples_per_step": total_samples,
                   "sample_definition": "one i16 PCM word that is decoded" + (" and then mixed" if mix else ""),
                   "layout": args.layout,
                   "l2": f"per step and GPU {shard.n * data_len / 1e6:.0f} MB of file images are read and {shard.samples * 2 / 1e6:.0f} MB of "
                         "samples written then re-read" + (": far larger than the 126 MB L2 (no flush needed)" if shard.n * data_len > 4e8 else
                                                           ": larger than the 126 MB L2 only in sum; no flush between steps — see the weak key for the > L2 per-GPU shard"),
                   "parallelism": f"strong scaling: the {n_files} files / voices sharded over {world} rank(s), file i -> rank i mod N" +
                                  (("; partial buses reduced tile by tile over peer memory (CUDA IPC / NVLink: " + ("inside the render kernel" if args.reduce == "p2p" else "two small kernels after the render") + "), no collective library"
                                    if peer is not None else "; one int32 all-reduce of the partial bus per step (NCCL)") if world > 1 and mix else "; no collective")},
        "roofline": roofline, "roofline_mix": roofline_mix, "roofline_step": roofline_step,
        "kernel_ms": {"decode": round(t["ms_decode_max"], 4), "mix": round(t["ms_mix_max"], 4),
                      "note": "max over ranks of each rank's per-step average"},
        "bus_check": bus_check, "weak": weak,
        "gpu_launches": int(t["launches"]), "clocks": clocks, "e2e": e2e, "e2e_parse_dropin": e2e_dropin,
    }
    if rank == 0 and not args.no_cpu and shard.n:
        out["cpu_baseline"] = cpu_baseline(shard.view, shard.image_len, shard.n, [gains[i] for i in shard.ids], threads=1,
                                           budget_s=args.cpu_seconds, mix=mix)
        out["cpu_fast_decode"] = cpu_fast_decode(shard.view, shard.image_len, shard.n)
    shard.close()
    if not args.no_configs:
        import bench_configs
        out["configs"] = bench_configs.run(ctx, rank, world, local, dist, peak, quick=args.quick_configs,
                                           reduce=args.reduce if args.reduce in ("p2p", "p2p2") else "p2p2")
    if rank == 0:
        print(json.dumps(out))
    if peer is not None:
        barrier(dist, local)
        peer.close()
    if dist is not None:
        dist.destroy_process_group()


def ctx_download(ctx, ptr, count, dtype=np.int16):
    out = np.empty(count, dtype=dtype)
    ctx.lib.blast_memcpy_d2h(ctx.h, out.ctypes.data, ptr, out.nbytes)
    ctx.sync()
    return out


# ------------------------------------------------------------------ CPU arms (oracle = checker, timed as the baseline)
def _cpu_decode_mix(view, image_len, idx, gains, mix):
    """faithful aiff::parse of files idx, then (mix) Conductor::coordinate over them as voices -> words"""
    import oracle
    L = oracle.lib()
    d = oracle.PcmDesc()
    words = 0
    bufs = []
    for i in idx:
        out = C.c_void_p()
        cnt = C.c_size_t()
        rc = L.orc_aiff_parse(view[i].ctypes.data, image_len, C.byref(d), C.byref(out), C.byref(cnt))
        assert rc == 0
        words += cnt.value
        bufs.append((out, cnt.value))
    if mix and bufs:
        arr = (oracle.Track * len(bufs))(*[oracle.Track(b.value, n, 2, 48000) for b, n in bufs])
        h = L.orc_conductor_new(2, 48000, arr, len(bufs))
        for k, i in enumerate(idx):
            L.orc_conductor_apply(h, C.byref(oracle.Command(kind=oracle.CMD_LOAD, idx=k, tempo=oracle.tempo_repr())))
            L.orc_conductor_apply(h, C.byref(oracle.Command(kind=oracle.CMD_START, idx_kind=oracle.IDX_VOICE, idx=k)))
            L.orc_conductor_set_voice(h, -1, k, None, None, C.byref(C.c_float(gains[i])), None)
        frames = bufs[0][1] // 2
        bus = np.empty(frames * 2, dtype=np.int16)
        L.orc_conductor_coordinate(h, frames, bus.ctypes.data)
        L.orc_conductor_free(h)
    for b, _ in bufs:
        L.orc_free(b)
    return words


def cpu_baseline(view, image_len, n_files, gains, threads: int, budget_s: float, mix: bool = True):
    """faithful CPU restatement (per-pair bounds-checked reads, Vec growth; frame->channel->voice scalar
    render loop) on a bounded sample; threads > 1 shard the files / voices (generous comparison)"""
    import oracle
    oracle.lib()
    # the cost per file grows with the batch (the frame -> channel -> voice loop strides over every track), so the sample
    # is sized from a probe batch and capped: the run must fit the budget
    probe = min(n_files, 8)
    t0 = time.perf_counter()
    _cpu_decode_mix(view, image_len, list(range(probe)), gains, mix)
    per_file = max(1e-4, (time.perf_counter() - t0) / probe)
    n = int(max(threads, min(n_files, 256 * threads, budget_s / (2.0 * per_file) * threads)))
    n = max(threads, n // threads * threads)
    n = min(n, n_files)
    idx = list(range(n))
    t0 = time.perf_counter()
    if threads == 1:
        words = _cpu_decode_mix(view, image_len, idx, gains, mix)
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(threads) as ex:
            words = sum(ex.map(lambda k: _cpu_decode_mix(view, image_len, idx[k::threads], gains, mix), range(threads)))
    dt = time.perf_counter() - t0
    what = "aiff::parse" + (" + Conductor::coordinate (each thread mixes its own voice shard)" if mix else "")
    return {"value": round(words / dt / 1e9, 4), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} of the C2 files ({words} i16 words, the same X128P bytes and gains as the GPU arm) through the faithful C++ "
                      f"restatement of {what} (oracle/blast_oracle.cpp, g++ -O2 -ffp-contract=off), {dt:.1f} s",
            "note": "CPU restatement of the reference, not the Rust binary (no rustc in the image)"}


def cpu_fast_decode(view, image_len, n_files, budget_s: float = 3.0):
    """the "good CPU" decode (bswap into a preallocated buffer, all host threads) on a bounded sample —
    decode only, there is no fast CPU mix; reported next to the faithful restatement, never as a gate"""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    L = oracle.lib()
    threads = os.cpu_count() or 1
    d = oracle.PcmDesc()
    assert L.orc_aiff_probe(view[0].ctypes.data, image_len, C.byref(d)) == 0
    words = L.orc_pcm_out_len(C.byref(d))
    n = min(n_files, 64 * threads)
    outs = [np.empty(words, dtype=np.int16) for _ in range(threads)]

    def work(k):
        for i in range(k, n, threads):
            assert L.orc_pcm_decode_fast(view[i].ctypes.data, image_len, C.byref(d), outs[k].ctypes.data) == 0

    t0, reps = time.perf_counter(), 0
    with ThreadPoolExecutor(threads) as ex:
        while time.perf_counter() - t0 < budget_s:
            list(ex.map(work, range(threads)))
            reps += 1
    dt = time.perf_counter() - t0
    return {"value": round(reps * n * words / dt / 1e9, 3), "unit": "Gwords/s decoded (no mix)", "cores": threads, "kind": "port",
            "sample": f"{reps} x {n} files through orc_pcm_decode_fast (byte swap into a preallocated buffer), {dt:.1f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import oracle
    n_files, data_len = args.files, args.data_len
    mix = not args.no_mix
    threads = os.cpu_count() or 1
    hdr = np.frombuffer(synth.aiff_header(data_len), dtype=np.uint8)
    image_len = len(hdr) + data_len
    sample_files = min(n_files, args.ref_files_per_thread * threads)
    view = np.empty((sample_files, image_len), dtype=np.uint8)
    view[:, :len(hdr)] = hdr
    for i in range(sample_files):                                       # the GPU arm's inputs: X128P::new(0xC2_0000 + i) bytes
        view[i, len(hdr):] = oracle.Rng(0xC20000 + i).fill_u64((data_len + 7) // 8).view(np.uint8)[:data_len]
    gains = c2_gains(n_files)
    res, times = None, []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        res = cpu_baseline(view, image_len, sample_files, gains, threads=threads, budget_s=1e9, mix=mix)
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
    words = sample_files * ((data_len + 1) // 2)
    value = words / (sum(times) / len(times)) / 1e9
    res["value"] = round(value, 4)
    workload = (f"C2: batch decode {n_files} x 24-bit BE 48 kHz stereo AIFF ({data_len} payload B each, X128P::new(0xC2_0000 + i) bytes), "
                "reference-exact byte-pair decode to i16" + (", then mix of the decoded tracks into one stereo S16 bus" if mix else ""))
    out = {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * sum(times) / len(times), 3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "i16", "data": "synthetic",
        "config": {"workload": workload, "sample_files_per_step": sample_files, "threads": threads,
                   "sample": f"each step decodes + mixes the first {sample_files} of the {n_files} files (same per-file work, "
                             "throughput-normalised subset of the GPU arm's batch) on all host threads"},
        "cpu_baseline": res,
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--files", type=int, default=synth.C2_FILES)
    ap.add_argument("--data-len", type=int, default=synth.C2_DATA_LEN)
    ap.add_argument("--layout", default="image", choices=["image", "payload"],
                    help="HBM-resident input: whole file images (payload at +54, misaligned) or 16B-aligned payloads")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-mix", action="store_true", help="decode only (no render/mix of the decoded tracks)")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the weak-scaling extra key")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configs (C1, C3, C4, C5, true 24-bit)")
    ap.add_argument("--quick-configs", action="store_true", help="other configs at reduced sizes (development)")
    ap.add_argument("--ref-files-per-thread", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--reduce", default="p2p2", choices=["p2p2", "p2p", "nccl"],
                    help="the mix reduction over peer memory as two small kernels after the render (p2p2, default: the fastest "
                         "on B200s), the same tile exchange inside the render kernel (p2p), or an NCCL all-reduce + finalize")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
