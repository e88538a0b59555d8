#!/usr/bin/env python
"""bench.py — BLAST hot path on B200: PCM decode+mix Gsamples/s (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, libblast_cuda.so)
  python bench.py --impl reference --gpus N ...            # CPU restatement of the reference, all host threads

Workload (config.workload): BASELINE config[1] — batch decode of 1,024 synthetic 24-bit big-endian
48 kHz stereo AIFF files (2,880,000 payload bytes each, reference-exact byte-pair decode into i16
words), followed by the mix of the decoded tracks where the render path is built.  One "step" = one
pass over the whole batch.  At N > 1 every rank owns its own 1,024-file shard (weak scaling,
no data-path collective for decode).

The JSON line follows the driver contract; `value` is device-resident throughput (CUDA events on the
launching stream), `e2e` the same metric through the host-buffer C-ABI call (pinned host file images in, S16
bus out, copies inside the timed region; decoded tracks stay in HBM), `e2e_parse_dropin` the variant that also
returns every AudioFile.samples Vec to the host like a literal aiff::parse() drop-in.
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


def dist_setup(n_gpus: int):
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


# ------------------------------------------------------------------ our arm
def run_ours(args):
    import audio_decoder_b200 as blast
    from audio_decoder_b200 import _lib, audio_processing as ap, blast_rand as br, file_parsing as fp

    rank, world, local, dist = dist_setup(args.gpus)
    torch = None
    if world > 1:
        import torch
        ctx = blast.Context(local, stream=torch.cuda.current_stream().cuda_stream)
    else:
        ctx = blast.Context(local)
    L = ctx.lib
    n_files, data_len = args.files, args.data_len
    hdr = np.frombuffer(synth.aiff_header(data_len), dtype=np.uint8)
    image_len = len(hdr) + data_len
    slot = (image_len + 255) // 256 * 256                      # file images at 256-byte aligned slots in HBM
    words_per_file = (data_len + 1) // 2
    frames_per_file = words_per_file // 2                      # stereo

    # ---- synthetic file images: pinned host slab (e2e input; the images lie back to back like an asset directory
    #      read into one buffer) + HBM slab (device-resident input, one 256-byte aligned slot per file)
    h_in = ctx.pinned(n_files * image_len)
    rng = np.random.default_rng(0xC20000 + rank)
    view = h_in.u8.reshape(n_files, image_len)
    view[:, :len(hdr)] = hdr
    chunk = 64
    for i in range(0, n_files, chunk):
        view[i:i + chunk, len(hdr):image_len] = rng.integers(0, 256, size=(min(chunk, n_files - i), data_len), dtype=np.uint8)
    d_in = ctx.alloc(n_files * slot)
    d_out = ctx.alloc(n_files * words_per_file * 2)
    for i in range(n_files):
        L.blast_memcpy_h2d(ctx.h, d_in.ptr + i * slot, h_in.ptr + i * image_len, image_len)
    ctx.sync()
    descs = [fp.probe("aiff", view[0, :image_len])] * n_files
    off = descs[0].data_off
    if args.layout == "payload":
        # payload-only layout: what blast_pcm_decode_batch stages (16-byte aligned payloads)
        d_pay = ctx.alloc(n_files * slot)
        for i in range(n_files):
            L.blast_memcpy_h2d(ctx.h, d_pay.ptr + i * slot, h_in.ptr + i * image_len + off, data_len)
        ctx.sync()
        d_src_base, src_extra = d_pay.ptr, 0
    else:
        d_src_base, src_extra = d_in.ptr, off
    jobs = [(d_src_base + i * slot + src_extra, d_out.ptr + i * words_per_file * 2, words_per_file, True)
            for i in range(n_files)]
    plan = fp.PcmPlan(ctx, jobs)
    samples_per_step = n_files * words_per_file
    alg_bytes_decode = 4 * samples_per_step                      # 2 B read + 2 B written per i16 word

    # ---- mix: every decoded file is one stereo voice (velocity 1, per-voice gain) on one stereo bus
    mix = not args.no_mix
    if mix:
        tracks, voices = [], []
        g = br.fill(ctx, 0xC2, 0, 1, n_files, 0, 100, ranged=False, checks=False)[0][0]
        for i in range(n_files):
            buf = blast.DevBuf.__new__(blast.DevBuf)
            buf.ctx, buf.ptr, buf.nbytes = ctx, d_out.ptr + i * words_per_file * 2, words_per_file * 2
            buf.free = lambda: None
            tracks.append(ap.Track(buf, words_per_file, 2, 48000))
            gain = float(np.float32((int(g[i]) >> 11) * 2.0 ** -53) * np.float32(2.0 ** -5))
            voices.append(ap.VoiceParams(i, True, 0.0, 1.0, gain))
        scene = ap.Scene(ctx, tracks, voices, 2)
        n_slots = frames_per_file * 2
        peer = None
        if world > 1 and args.reduce == "p2p":
            # the one exchange step over peer memory: partial buses mapped into rank 0, ONE reduce + finalize kernel
            from audio_decoder_b200 import distributed as bd
            try:
                peer = bd.PeerBus(ctx, n_slots, rank, world, mode=args.peer_mode)
                part_ptr = peer.part.ptr
            except RuntimeError as e:                              # raised on EVERY rank or on none
                if rank == 0:
                    print(f"bench.py: {e}; using the NCCL all-reduce instead", file=sys.stderr)
                peer = None
                args.reduce = "nccl"
        if world > 1 and peer is None:
            t_part = torch.empty(n_slots, dtype=torch.int32, device=f"cuda:{local}")
            part_ptr = t_part.data_ptr()
        elif peer is None:
            d_part = ctx.alloc(4 * n_slots)
            part_ptr = d_part.ptr
        d_bus = peer.bus if peer is not None else ctx.alloc(2 * n_slots)
        alg_bytes_mix = 4 * (frames_per_file - 1) * n_files + 2 * n_slots   # source frames touched + S16 bus
    n_ev = 3 if mix else 2

    def mix_step():
        if peer is not None:
            peer.wait_ack()                                        # rank 0 has consumed the previous partial bus
            scene.restore_dev()
            scene.render_partial_dev(frames_per_file, part_ptr)
            peer.reduce()                                          # signal, then ONE kernel: wait + reduce my slice + wrap + store to rank 0
            return
        scene.restore_dev()
        scene.render_partial_dev(frames_per_file, part_ptr)
        if world > 1:
            dist.all_reduce(t_part, op=dist.ReduceOp.SUM)          # NCCL variant: int32 partial buses
        ap.finalize_bus(ctx, part_ptr, d_bus.ptr, n_slots)

    def step(evs=None):
        if evs:
            evs[0].record()
        plan.run()
        if evs:
            evs[1].record()
        if mix:
            mix_step()
            if evs:
                evs[2].record()

    for _ in range(args.warmup):
        step()
    ctx.sync()
    if mix:
        scene.check()
    barrier(dist, local)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count
    evs = [[ctx.event() for _ in range(n_ev)] for _ in range(args.steps)]
    e_end = ctx.event()
    for k in range(args.steps):
        step(evs[k])
    e_end.record()
    ms_total = evs[0][0].elapsed_ms(e_end)
    ctx.sync()
    barrier(dist, local)
    launches = ctx.launch_count - launches0
    clocks = None
    ms_decode = sum(e[0].elapsed_ms(e[1]) for e in evs) / args.steps
    ms_mix = sum(e[1].elapsed_ms(e[2]) for e in evs) / args.steps if mix else 0.0
    ms_total = max_over_ranks(dist, local, ms_total)
    ms_step = ms_total / args.steps
    value = world * samples_per_step / (ms_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (decode moves 2x the bytes of the mix), per-launch average
    peak, peak_src = measured_peaks()
    achieved = alg_bytes_decode / (ms_decode * 1e-3) / 1e9
    roofline = {"kernel": "pcm16_decode_batch", "bound": "hbm", "achieved": round(achieved, 1), "peak": peak,
                "unit": "GB/s", "frac": round(achieved / peak, 4), "frac_of_nominal_8000": round(achieved / 8000.0, 4),
                "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes_decode,
                "ms_per_launch": round(ms_decode, 4)}
    traffic = {}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            traffic = json.load(open(tr))
        except Exception:
            traffic = {}
    roofline["traffic"] = traffic.get("pcm16_decode_batch")
    roofline_mix = None
    if mix:
        ach = alg_bytes_mix / (ms_mix * 1e-3) / 1e9
        roofline_mix = {"kernel": "voice_position_scan + voice_render_mix_tma + " +
                                  ("bus_finalize" if world == 1 else "bus_reduce_peers (peer memory)" if peer is not None
                                   else "NCCL all-reduce(int32) + bus_finalize"),
                        "bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                        "frac": round(ach / peak, 4), "algorithmic_bytes_per_step": alg_bytes_mix,
                        "ms_per_step": round(ms_mix, 4), "traffic": traffic.get("voice_render_mix_tma_c2")}

    # the whole step against the same roofline (north star: decode + mix at >= 70 % of the HBM roofline per GPU)
    alg_step = alg_bytes_decode + (alg_bytes_mix if mix else 0)
    roofline_step = {"bound": "hbm", "achieved": round(alg_step / (ms_step * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(alg_step / (ms_step * 1e-3) / 1e9 / peak, 4), "algorithmic_bytes_per_step": alg_step,
                     "note": "decode + mix of one rank, max-over-ranks step time"}

    # ---- e2e: the same step through the host-buffer C ABI.  Pinned host file images -> blast_pcm_decode_batch
    #      (H2D inside) -> render of the decoded tracks -> S16 bus copied back to the host.
    #      "e2e":        AudioFile.samples stay in HBM (their only consumer is the render kernel; SURVEY §8 b)
    #      "e2e_parse_dropin": additionally every AudioFile.samples Vec is delivered to the host, as a literal
    #                    aiff::parse() drop-in must (doubles the PCIe traffic)
    e2e = e2e_dropin = None
    if not args.no_e2e:
        h_out = ctx.pinned(n_files * words_per_file * 2)
        h_bus = ctx.pinned(2 * frames_per_file * 2) if mix else None
        files = (C.c_void_p * n_files)(*[h_in.ptr + i * image_len for i in range(n_files)])
        lens = (C.c_size_t * n_files)(*([image_len] * n_files))
        dd = (_lib.PcmDesc * n_files)(*descs)
        host_out = (C.c_void_p * n_files)(*[h_out.ptr + i * words_per_file * 2 for i in range(n_files)])
        dev_out = (C.c_void_p * n_files)(*[d_out.ptr + i * words_per_file * 2 for i in range(n_files)])
        pcie = {}
        pp = os.path.join(ROOT, "profiles", "r01_pcie_probe.json")
        if os.path.exists(pp):
            pcie = json.load(open(pp))

        def e2e_step(to_host):
            rc = L.blast_pcm_decode_batch(ctx.h, n_files, files, lens, dd, host_out if to_host else None, dev_out)
            if rc != 0:
                raise RuntimeError(L.blast_last_error().decode())
            if mix:
                mix_step()
                if peer is None or rank == 0:                      # over peer memory the bus exists on rank 0 only
                    L.blast_memcpy_d2h(ctx.h, h_bus.ptr, d_bus.ptr, 2 * n_slots)
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
            h2d = n_files * data_len
            d2h = (n_files * words_per_file * 2 if to_host else 0) + (2 * n_slots if mix else 0)
            r = {"value": round(world * samples_per_step / dt / 1e9, 3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                 "d2h_bytes_per_step": d2h, "ms_per_step": round(dt * 1e3, 3),
                 "pcie_GBps": round(max(h2d, d2h) / dt / 1e9, 2)}
            if pcie:
                r["pcie_peak_GBps"] = pcie.get("bidir_each_GBps" if to_host else "h2d_GBps")
                r["pcie_frac"] = round(r["pcie_GBps"] / r["pcie_peak_GBps"], 3)
                r["pcie_peak_source"] = "tools/pcie_probe.py on this pool's B200 box (profiles/r01_pcie_probe.json)"
            return r

        e2e = e2e_run(False)
        e2e["api"] = ("blast_pcm_decode_batch (pinned host file images in, decoded tracks kept in HBM)" +
                      (" + blast_scene_render_dev + blast_bus_finalize_dev + S16 bus D2H" if mix else ""))
        e2e_dropin = e2e_run(True)
        e2e_dropin["api"] = e2e["api"].replace("decoded tracks kept in HBM", "host AudioFile.samples out AND tracks kept in HBM")
        # spot-check the e2e result against numpy (not timed)
        got = h_out.view(np.int16, words_per_file, 0)
        assert np.array_equal(got, view[0, off:off + data_len].view(">i2").astype(np.int16)), "e2e output mismatch"
    if rank == 0:
        clocks = sampler.stop()

    workload = (f"C2: batch decode {n_files} x 24-bit BE 48 kHz stereo AIFF ({data_len} payload B each), reference-exact "
                "byte-pair decode to i16" + (f", then mix of the {n_files} decoded tracks (stereo voices, velocity 1, "
                                             "per-voice gain) into one stereo S16 bus" if mix else ""))
    out = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "i16", "data": "synthetic",
        "config": {"workload": workload, "files_per_gpu": n_files, "samples_per_step_per_gpu": samples_per_step,
                   "sample_definition": "one i16 PCM word that is decoded" + (" and then mixed" if mix else ""),
                   "layout": args.layout,
                   "l2": f"per step {n_files * data_len / 1e6:.0f} MB of file images are read and {samples_per_step * 2 / 1e6:.0f} MB of "
                         "samples written then re-read: far larger than the 126 MB L2 (no flush needed)",
                   "parallelism": f"files / voices sharded over {world} rank(s)" +
                                  ((f"; partial buses reduced + finalized over peer memory (CUDA IPC / NVLink, mode {args.peer_mode}: one fused kernel), no collective library"
                                    if peer is not None else "; one int32 all-reduce of the partial bus per step (NCCL)") if world > 1 and mix else "; no collective")},
        "roofline": roofline, "roofline_mix": roofline_mix, "roofline_step": roofline_step,
        "kernel_ms": {"decode": round(ms_decode, 4), "mix": round(ms_mix, 4)},
        "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "e2e_parse_dropin": e2e_dropin,
    }
    if rank == 0 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(view, image_len, n_files, threads=1, budget_s=args.cpu_seconds, mix=mix)
        out["cpu_fast_decode"] = cpu_fast_decode(view, image_len, n_files)
    if rank == 0:
        print(json.dumps(out))
    plan.close()
    if mix:
        if peer is not None:
            barrier(dist, local)
            peer.close()
        scene.close()
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------ CPU arms (oracle = checker, timed as the baseline)
def _cpu_decode_mix(view, image_len, idx, mix):
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
        for k in range(len(bufs)):
            L.orc_conductor_apply(h, C.byref(oracle.Command(kind=oracle.CMD_LOAD, idx=k, tempo=oracle.tempo_repr())))
            L.orc_conductor_apply(h, C.byref(oracle.Command(kind=oracle.CMD_START, idx_kind=oracle.IDX_VOICE, idx=k)))
            L.orc_conductor_set_voice(h, -1, k, None, None, C.byref(C.c_float(0.01)), None)
        frames = bufs[0][1] // 2
        bus = np.empty(frames * 2, dtype=np.int16)
        L.orc_conductor_coordinate(h, frames, bus.ctypes.data)
        L.orc_conductor_free(h)
    for b, _ in bufs:
        L.orc_free(b)
    return words


def cpu_baseline(view, image_len, n_files, threads: int, budget_s: float, mix: bool = True):
    """faithful CPU restatement (per-pair bounds-checked reads, Vec growth; frame->channel->voice scalar
    render loop) on a bounded sample; threads > 1 shard the files / voices (generous comparison)"""
    import oracle
    oracle.lib()
    t0 = time.perf_counter()
    _cpu_decode_mix(view, image_len, [0], mix)
    per_file = max(1e-4, time.perf_counter() - t0)
    n = int(max(threads, min(n_files, budget_s / per_file * threads)))
    n = max(threads, n // threads * threads)
    idx = list(range(n))
    t0 = time.perf_counter()
    if threads == 1:
        words = _cpu_decode_mix(view, image_len, idx, mix)
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(threads) as ex:
            words = sum(ex.map(lambda k: _cpu_decode_mix(view, image_len, idx[k::threads], mix), range(threads)))
    dt = time.perf_counter() - t0
    what = "aiff::parse" + (" + Conductor::coordinate (each thread mixes its own voice shard)" if mix else "")
    return {"value": round(words / dt / 1e9, 4), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} of {n_files} files ({words} i16 words) through the faithful C++ restatement of {what} "
                      f"(oracle/blast_oracle.cpp, g++ -O2 -ffp-contract=off), {dt:.1f} s",
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
    n_files, data_len = args.files, args.data_len
    mix = not args.no_mix
    threads = os.cpu_count() or 1
    hdr = np.frombuffer(synth.aiff_header(data_len), dtype=np.uint8)
    image_len = len(hdr) + data_len
    sample_files = min(n_files, args.ref_files_per_thread * threads)
    view = np.empty((sample_files, image_len), dtype=np.uint8)
    rng = np.random.default_rng(0xC20000)
    view[:, :len(hdr)] = hdr
    view[:, len(hdr):] = rng.integers(0, 256, size=(sample_files, data_len), dtype=np.uint8)
    res, times = None, []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        res = cpu_baseline(view, image_len, sample_files, threads=threads, budget_s=1e9, mix=mix)
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
    words = sample_files * ((data_len + 1) // 2)
    value = words / (sum(times) / len(times)) / 1e9
    res["value"] = round(value, 4)
    workload = (f"C2: batch decode {n_files} x 24-bit BE 48 kHz stereo AIFF ({data_len} payload B each), reference-exact "
                "byte-pair decode to i16" + (", then mix of the decoded tracks into one stereo S16 bus" if mix else ""))
    out = {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * sum(times) / len(times), 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i16", "data": "synthetic",
        "config": {"workload": workload, "sample_files_per_step": sample_files, "threads": threads},
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
    ap.add_argument("--ref-files-per-thread", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--peer-mode", default="root", choices=["root", "scatter"],
                    help="p2p reduction: the root pulls every bus (no lock step) / every rank reduces its 1/N slice")
    ap.add_argument("--reduce", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: the mix reduction over peer memory (one fused kernel on rank 0) or as an NCCL all-reduce")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
