#!/usr/bin/env python
"""C5: MPEG frame-sync scan over a 16 GiB synthetic stream (device-resident), CUDA-event timing."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audio_decoder_b200 as blast  # noqa: E402
import synth  # noqa: E402

if __name__ == "__main__":
    a = argparse.ArgumentParser()
    a.add_argument("--gib", type=int, default=16)
    a.add_argument("--iters", type=int, default=3)
    args = a.parse_args()
    with blast.Context(0) as ctx:
        L = ctx.lib
        block = synth.mp3_like(0xC5, (1 << 30) // 418 - 4)             # ~1 GiB block, zero tail
        block = np.concatenate([block, np.zeros((1 << 30) - block.size, np.uint8)])
        n = args.gib << 30
        d = ctx.alloc(n)
        h = ctx.pinned(block.size)
        h.u8[:] = block
        for k in range(args.gib):
            L.blast_memcpy_h2d(ctx.h, d.ptr + (k << 30), h.ptr, block.size)
        ctx.sync()
        cap = n // 256
        d_pos, d_hdr = ctx.alloc(8 * cap), ctx.alloc(4 * cap)
        cnt = C.c_uint64()
        times = []
        for it in range(args.iters + 1):
            e0 = ctx.event().record()
            rc = L.blast_mpeg_scan_dev(ctx.h, d.ptr, n, d_pos.ptr, d_hdr.ptr, cap, C.byref(cnt))
            e1 = ctx.event().record()
            assert rc == 0, L.blast_last_error()
            if it:
                times.append(e0.elapsed_ms(e1))
        ms = float(np.median(times))
        res = {"scan": {"gib": args.gib, "ms": round(ms, 3), "GBps_scanned": round(n / ms / 1e6, 1), "candidates": cnt.value,
                        "alg_GBps": round((n + 12 * cnt.value) / ms / 1e6, 1)}}
        noff, ncand, ref = C.c_uint64(), C.c_uint64(), C.c_uint32()
        d_off = ctx.alloc(8 * cap)
        its = []
        for it in range(args.iters + 1):                      # the first call grows the context scratch (untimed)
            e0 = ctx.event().record()
            rc = L.blast_mpeg_index_dev(ctx.h, d.ptr, n, 1, d_off.ptr, cap, C.byref(noff), C.byref(ref), C.byref(ncand))
            e1 = ctx.event().record()
            assert rc == 0, L.blast_last_error()
            if it:
                its.append(e0.elapsed_ms(e1))
        res["index"] = {"ms": round(float(np.median(its)), 3), "offsets": noff.value, "ref_header": hex(ref.value),
                        "candidates": ncand.value, "includes": "scan + header vote + classify + ordered compaction"}
        offs = d_off.download(np.uint64, min(noff.value, 1 << 20))
        res["index"]["sorted_prefix"] = bool(np.all(np.diff(offs.astype(np.int64)) >= 0))
        print(json.dumps(res, indent=1))
