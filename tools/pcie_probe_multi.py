#!/usr/bin/env python
"""What limits the host-buffer (e2e) path when several GPUs ingest at once: run under torchrun with one rank per GPU.
For k = 1, 2, 4, ... world concurrent ranks (the others idle) every active rank copies a pinned 1 GiB host buffer to
its GPU (H2D), back (D2H), and both at once; per-rank and aggregate GB/s.  Variants of where the pinned buffer lives:
"default" (first touch by this process) and, when libnuma's `numactl` is present, the buffer's NUMA node as reported by
/proc; the topology (`nvidia-smi topo -m`, CPU affinity of every GPU) is recorded with the table.  torch is plumbing
here (process group, pinned buffers, streams)."""
import json
import os
import subprocess
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
h_a.fill_(rank + 1)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_a, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_b.copy_(d_b, non_blocking=True)


def both():
    h2d(); d2h()


def run(fn, active, reps=4):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        dist.barrier(device_ids=[local])
        t0 = time.perf_counter()
        if active:
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 2
        best = min(best, dt)
    t = torch.tensor([n / best / 1e9 if active else 0.0], dtype=torch.float64, device="cuda")
    all_t = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(all_t, t)
    return [round(float(x), 2) for x in all_t]


res = {"world": world, "bytes_per_copy": n, "rows": []}
k = 1
while k <= world:
    active = rank < k
    row = {"concurrent_ranks": k}
    for name, fn in (("h2d", h2d), ("d2h", d2h), ("bidir_each", both)):
        per = run(fn, active)
        row[name + "_per_rank_GBps"] = per[:k]
        row[name + "_sum_GBps"] = round(sum(per[:k]), 1)
    res["rows"].append(row)
    k *= 2
if rank == 0:
    def sh(cmd):
        try:
            return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout
        except Exception as e:  # noqa: BLE001
            return str(e)
    res["topo"] = sh("nvidia-smi topo -m")
    res["lscpu"] = [ln for ln in sh("lscpu").splitlines() if any(s in ln for s in ("Model name", "Socket", "NUMA", "CPU(s):", "Thread"))]
    res["numactl"] = sh("numactl --hardware 2>&1 | head -20")
    res["meminfo"] = [ln for ln in sh("cat /proc/meminfo").splitlines()[:3]]
    print(json.dumps(res))
dist.barrier(device_ids=[local])
dist.destroy_process_group()
