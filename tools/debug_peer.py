#!/usr/bin/env python
"""two peer buses on one GPU driven from one thread, step by step, with flag dumps (development)"""
import ctypes as C
import time
import os
import sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("BLAST_PEER_TIMEOUT_MS", "3000")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import audio_decoder_b200 as blast
from audio_decoder_b200 import _lib, audio_processing as ap
from audio_decoder_b200.errors import check

L = _lib.load()
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctxs = [blast.Context(0) for _ in range(world)]
pbs = []
for r, c in enumerate(ctxs):
    p = C.c_void_p()
    check(L.blast_peer_bus_create(c.h, max(2 * frames, 8), r, world, 0, C.byref(p)))
    pbs.append(p.value)
arr = (C.c_void_p * world)(*pbs)
check(L.blast_peer_bus_connect_local(arr, world))
rng = np.random.default_rng(1)
clip = rng.integers(-3000, 3000, size=(frames + 8) * 2).astype(np.int16)
scenes = []
for r, c in enumerate(ctxs):
    t = ap.Track.from_host(c, clip, 2)
    sc = ap.Scene(c, [t], [ap.VoiceParams(0, True, 0.0, 1.0, 1.0)], 2)
    check(L.blast_scene_reserve(c.h, sc.h, frames))
    scenes.append((t, sc))


def dump(tag):
    for r, c in enumerate(ctxs):
        out = (C.c_uint32 * 64)()
        L.blast_peer_bus_flags(c.h, pbs[r], out, 64)
        print(tag, "rank", r, list(out[:2 + 6 * world]), "err/red", list(out[2 + 6 * world:4 + 6 * world]), flush=True)


dump("created")
for step in range(2):
    for r, c in enumerate(ctxs):
        check(L.blast_scene_render_reduce_dev(c.h, scenes[r][1].h, frames, pbs[r]))
        print("enqueued rank", r, flush=True)
        if step == 0 and r == 0:
            time.sleep(0.2)
            dump("rank 0 alone")
    time.sleep(0.05)
    dump(f"step {step} after 0.05 s")
    t0 = time.perf_counter()
    check(L.blast_peer_bus_wait_dev(ctxs[0].h, pbs[0]))
    ctxs[0].sync()
    print("wait+sync took", round(time.perf_counter() - t0, 3), "s", flush=True)
    bus = np.zeros(frames * 2, np.int16)
    check(L.blast_memcpy_d2h(ctxs[0].h, bus.ctypes.data, L.blast_peer_bus_bus(pbs[0]), bus.nbytes))
    for r, c in enumerate(ctxs):
        rc = L.blast_peer_bus_check(c.h, pbs[r])
        print("check rank", r, rc, L.blast_last_error().decode() if rc else "", flush=True)
    want = (clip[:frames * 2].astype(np.int32) * world).astype(np.int16)
    print("bus ok:", np.array_equal(bus, want), flush=True)
    dump(f"step {step} end")
