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
launching stream), `e2e` the same metric through the host-buffer C-ABI call (pinned host file images in,
host AudioFile.samples out, copies inside the timed region).
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
    from audio_decoder_b200 import _lib, file_parsing as fp

    rank, world, local, dist = dist_setup(args.gpus)
    ctx = blast.Context(local)
    L = ctx.lib
    n_files, data_len = args.files, args.data_len
    hdr = np.frombuffer(synth.aiff_header(data_len), dtype=np.uint8)
    image_len = len(hdr) + data_len
    slot = (image_len + 255) // 256 * 256                      # file images at 256-byte aligned slots in HBM
    words_per_file = (data_len + 1) // 2

    # ---- synthetic file images: pinned host slab (e2e input) + HBM slab (device-resident input)
    h_in = ctx.pinned(n_files * slot)
    rng = np.random.default_rng(0xC20000 + rank)
    view = h_in.u8.reshape(n_files, slot)
    view[:, :len(hdr)] = hdr
    chunk = 64
    for i in range(0, n_files, chunk):
        view[i:i + chunk, len(hdr):image_len] = rng.integers(0, 256, size=(min(chunk, n_files - i), data_len), dtype=np.uint8)
    d_in = ctx.alloc(n_files * slot)
    d_out = ctx.alloc(n_files * words_per_file * 2)
    L.blast_memcpy_h2d(ctx.h, d_in.ptr, h_in.ptr, n_files * slot)
    ctx.sync()
    descs = [fp.probe("aiff", view[0, :image_len])] * n_files
    off = descs[0].data_off
    src_extra = off if args.layout == "image" else 0
    if args.layout == "payload":
        # payload-only layout: what blast_pcm_decode_batch stages (16-byte aligned payloads)
        d_pay = ctx.alloc(n_files * slot)
        for i in range(n_files):
            L.blast_memcpy_h2d(ctx.h, d_pay.ptr + i * slot, h_in.ptr + i * slot + off, data_len)
        ctx.sync()
        d_src_base = d_pay.ptr
    else:
        d_src_base = d_in.ptr
    jobs = [(d_src_base + i * slot + src_extra, d_out.ptr + i * words_per_file * 2, words_per_file, True)
            for i in range(n_files)]
    plan = fp.PcmPlan(ctx, jobs)
    samples_per_step = n_files * words_per_file
    alg_bytes_decode = 4 * samples_per_step                      # 2 B read + 2 B written per i16 word

    def step():
        plan.run()

    for _ in range(args.warmup):
        step()
    ctx.sync()
    barrier(dist, local)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count
    e0, e1 = ctx.event(), ctx.event()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    ms_total = e0.elapsed_ms(e1)
    ctx.sync()
    barrier(dist, local)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(dist, local, ms_total)
    ms_step = ms_total / args.steps
    value = world * samples_per_step / (ms_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (decode): per-launch average over the timed region
    peak, peak_src = measured_peaks()
    kern_ms = ms_step                                             # one launch per step
    achieved = alg_bytes_decode / (kern_ms * 1e-3) / 1e9
    roofline = {"kernel": "pcm16_decode_batch", "bound": "hbm", "achieved": round(achieved, 1), "peak": peak,
                "unit": "GB/s", "frac": round(achieved / peak, 4), "frac_of_nominal_8000": round(achieved / 8000.0, 4),
                "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes_decode}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            roofline["traffic"] = json.load(open(tr)).get("pcm16_decode_batch")
        except Exception:
            pass

    # ---- e2e: host file images (pinned) -> blast_pcm_decode_batch -> host AudioFile.samples (pinned)
    e2e = None
    if not args.no_e2e:
        h_out = ctx.pinned(n_files * words_per_file * 2)
        files = (C.c_void_p * n_files)(*[h_in.ptr + i * slot for i in range(n_files)])
        lens = (C.c_size_t * n_files)(*([image_len] * n_files))
        dd = (_lib.PcmDesc * n_files)(*descs)
        host_out = (C.c_void_p * n_files)(*[h_out.ptr + i * words_per_file * 2 for i in range(n_files)])
        dev_out = (C.c_void_p * n_files)(*[d_out.ptr + i * words_per_file * 2 for i in range(n_files)])

        def e2e_step():
            rc = L.blast_pcm_decode_batch(ctx.h, n_files, files, lens, dd, host_out, dev_out)
            if rc != 0:
                raise RuntimeError(L.blast_last_error().decode())

        for _ in range(max(1, min(2, args.warmup))):
            e2e_step()
        barrier(dist, local)
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        ctx.sync()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        dt = max_over_ranks(dist, local, dt)
        # spot-check the e2e result against numpy (not timed)
        got = h_out.view(np.int16, words_per_file, 0)
        assert np.array_equal(got, view[0, off:off + data_len].view(">i2").astype(np.int16)), "e2e output mismatch"
        e2e = {"value": round(world * samples_per_step / dt / 1e9, 3), "unit": UNIT,
               "h2d_bytes_per_step": n_files * data_len, "d2h_bytes_per_step": n_files * words_per_file * 2,
               "ms_per_step": round(dt * 1e3, 3), "api": "blast_pcm_decode_batch (host images in, host samples out)"}

    out = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "i16", "data": "synthetic",
        "config": {"workload": f"C2: batch decode {n_files} x 24-bit BE 48 kHz stereo AIFF ({data_len} payload B each), "
                               "reference-exact byte-pair decode to i16",
                   "files_per_gpu": n_files, "samples_per_step_per_gpu": samples_per_step, "layout": args.layout,
                   "l2": f"input {n_files * data_len / 1e6:.0f} MB + output per step, far larger than the 126 MB L2 (no flush needed)",
                   "parallelism": f"files sharded, {world} rank(s), no collective"},
        "roofline": roofline, "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e,
    }
    if rank == 0 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(view, image_len, n_files, threads=1, budget_s=args.cpu_seconds)
    if rank == 0:
        print(json.dumps(out))
    plan.close()
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------ CPU arms (oracle = checker, timed as the baseline)
def _cpu_parse_files(view, image_len, idx):
    import oracle
    L = oracle.lib()
    d = oracle.PcmDesc()
    out = C.c_void_p()
    cnt = C.c_size_t()
    words = 0
    for i in idx:
        rc = L.orc_aiff_parse(view[i].ctypes.data, image_len, C.byref(d), C.byref(out), C.byref(cnt))
        assert rc == 0
        words += cnt.value
        L.orc_free(out)
    return words


def cpu_baseline(view, image_len, n_files, threads: int, budget_s: float):
    """faithful CPU restatement (per-pair bounds-checked reads, Vec growth) on a bounded sample"""
    import oracle
    oracle.lib()
    t0 = time.perf_counter()
    w = _cpu_parse_files(view, image_len, [0])
    per_file = max(1e-4, time.perf_counter() - t0)
    n = int(max(threads, min(n_files, budget_s / per_file * threads)))
    n = n // threads * threads
    idx = list(range(n))
    t0 = time.perf_counter()
    if threads == 1:
        words = _cpu_parse_files(view, image_len, idx)
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(threads) as ex:
            words = sum(ex.map(lambda k: _cpu_parse_files(view, image_len, idx[k::threads]), range(threads)))
    dt = time.perf_counter() - t0
    return {"value": round(words / dt / 1e9, 4), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} of {n_files} files ({words} i16 words) through the faithful C++ restatement of aiff::parse "
                      f"(oracle/blast_oracle.cpp, g++ -O2), {dt:.1f} s",
            "note": "CPU restatement of the reference, not the Rust binary (no rustc in the image)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    n_files, data_len = args.files, args.data_len
    threads = os.cpu_count() or 1
    hdr = np.frombuffer(synth.aiff_header(data_len), dtype=np.uint8)
    image_len = len(hdr) + data_len
    sample_files = min(n_files, max(threads, 4 * threads))
    view = np.empty((sample_files, image_len), dtype=np.uint8)
    rng = np.random.default_rng(0xC20000)
    view[:, :len(hdr)] = hdr
    view[:, len(hdr):] = rng.integers(0, 256, size=(sample_files, data_len), dtype=np.uint8)
    res = None
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        res = cpu_baseline(view, image_len, sample_files, threads=threads, budget_s=1e9)
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
    value = res["value"]
    words = sample_files * ((data_len + 1) // 2)
    value = words / (sum(times) / len(times)) / 1e9
    res["value"] = round(value, 4)
    out = {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * sum(times) / len(times), 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i16", "data": "synthetic",
        "config": {"workload": f"C2: batch decode {n_files} x 24-bit BE 48 kHz stereo AIFF ({data_len} payload B each), "
                               "reference-exact byte-pair decode to i16",
                   "sample_files_per_step": sample_files, "threads": threads},
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
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
