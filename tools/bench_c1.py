#!/usr/bin/env python
"""C1: decode ONE synthetic 10-minute 16-bit stereo 44.1 kHz WAV (105,840,044 bytes) via file_parsing::wav.
Device-resident: four copies of the file image are rotated (423 MB > the 126 MB L2) and decoded one per launch;
e2e: blast_pcm_decode_batch on the pinned host image, host AudioFile.samples out.  (The CPU side of the comparison is
tests/checks/c1_cpu_baseline.py: the oracle is test infrastructure.)"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audio_decoder_b200 as blast  # noqa: E402
from audio_decoder_b200 import _lib, file_parsing as fp  # noqa: E402
import synth  # noqa: E402

if __name__ == "__main__":
    with blast.Context(0) as ctx:
        L = ctx.lib
        img = synth.wav_image(0xC1, synth.C1_DATA_LEN)
        d = fp.probe("wav", img)
        words = L.blast_pcm_out_len(C.byref(d))
        h = ctx.pinned(img.size)
        h.u8[:] = img
        n_rot = 4
        slot = (img.size + 255) // 256 * 256
        d_in = ctx.alloc(n_rot * slot)
        d_out = ctx.alloc(n_rot * words * 2)
        for k in range(n_rot):
            L.blast_memcpy_h2d(ctx.h, d_in.ptr + k * slot, h.ptr, img.size)
        ctx.sync()
        plans = [fp.PcmPlan(ctx, [(d_in.ptr + k * slot + d.data_off, d_out.ptr + k * words * 2, words, False)]) for k in range(n_rot)]
        for p in plans:
            p.run()
        ctx.sync()
        iters = 40
        e0 = ctx.event().record()
        for i in range(iters):
            plans[i % n_rot].run()
        e1 = ctx.event().record()
        ms = e0.elapsed_ms(e1) / iters
        got = d_out.download(np.int16, words)
        assert np.array_equal(got, img[44:].view("<i2")), "C1 decode mismatch"
        res = {"workload": "C1: one 10-min 16-bit stereo 44.1 kHz WAV, 52,920,000 samples",
               "device_resident": {"ms": round(ms, 4), "gsamples_per_s": round(words / ms / 1e6, 1),
                                   "GBps": round(4 * words / ms / 1e6, 1), "l2": "4 rotating file images (423 MB)"}}
        # e2e through the parse() drop-in: host image in, host samples out
        out = ctx.pinned(words * 2)
        files = (C.c_void_p * 1)(h.ptr)
        lens = (C.c_size_t * 1)(img.size)
        dd = (_lib.PcmDesc * 1)(d)
        host_out = (C.c_void_p * 1)(out.ptr)
        for _ in range(2):
            assert L.blast_pcm_decode_batch(ctx.h, 1, files, lens, dd, host_out, None) == 0
        t0 = time.perf_counter()
        for _ in range(5):
            assert L.blast_pcm_decode_batch(ctx.h, 1, files, lens, dd, host_out, None) == 0
        dt = (time.perf_counter() - t0) / 5
        assert np.array_equal(out.view(np.int16, words), img[44:].view("<i2"))
        res["e2e_parse_dropin"] = {"ms": round(dt * 1e3, 3), "gsamples_per_s": round(words / dt / 1e9, 2),
                                   "pcie_GBps_each_way": round(2 * words / dt / 1e9, 1)}
        print(json.dumps(res, indent=1))
