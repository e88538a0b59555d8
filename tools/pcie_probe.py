#!/usr/bin/env python
"""PCIe ceiling of the box (the e2e roofline): pinned H2D, D2H and both at once, 1 GiB each, CUDA events.
torch is used here as plumbing only (two streams + pinned buffers)."""
import json
import torch

n = 1 << 30
h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        s1.synchronize(); s2.synchronize()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_a, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_b.copy_(d_b, non_blocking=True)


def both():
    h2d(); d2h()


import time
def wall(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3

res = {"h2d_GBps": round(n / wall(h2d) / 1e6, 2), "d2h_GBps": round(n / wall(d2h) / 1e6, 2)}
ms = wall(both)
res["bidir_each_GBps"] = round(n / ms / 1e6, 2)
res["bidir_total_GBps"] = round(2 * n / ms / 1e6, 2)
print(json.dumps(res))
