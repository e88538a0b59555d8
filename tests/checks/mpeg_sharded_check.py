#!/usr/bin/env python
"""Run under torchrun (one rank per GPU): the frame-offset index of ONE stream cut into byte ranges over the ranks
(audio_decoder_b200.distributed.ShardedMpegIndex) against the CPU oracle on rank 0, then a timing of the sharded scan
on a large stream (GiB per rank given by --gib)."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import audio_decoder_b200 as blast  # noqa: E402
from audio_decoder_b200 import distributed as bd, file_parsing as fp  # noqa: E402
import synth  # noqa: E402

if __name__ == "__main__":
    a = argparse.ArgumentParser()
    a.add_argument("--gib", type=int, default=4)
    a.add_argument("--quick", action="store_true", help="parity only (the pytest -m gpu entry)")
    args = a.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = blast.Context(local, stream=torch.cuda.current_stream().cuda_stream)
    res = {"world": world}
    # ---- parity: every rank builds the same stream, keeps its range (+ halo) on its GPU
    buf = synth.mp3_like(0xC5, (48 << 20) // 418)
    start, own, halo = bd.mpeg_plan_ranges(buf.size, world)[rank]
    d = ctx.to_device(buf[start:start + own + halo])
    for compat in (True, False):
        got = bd.ShardedMpegIndex(ctx, rank, world).run(d.ptr, buf.size, reference_compat=compat)
        offs = got["d_offsets"].download(np.uint64, got["n_offsets"])
        t = torch.from_numpy(offs.astype(np.int64)).cuda()
        sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([t.numel()], dtype=torch.int64, device="cuda"))
        mx = int(max(int(s) for s in sizes))
        pad = torch.zeros(mx, dtype=torch.int64, device="cuda")
        pad[:t.numel()] = t
        allp = [torch.zeros(mx, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(allp, pad)
        if rank == 0:
            import oracle
            exp = oracle.mpeg_parse(buf, reference_compat=compat, want_payload=False)
            cat = np.concatenate([p[:int(s)].cpu().numpy() for p, s in zip(allp, sizes)]).astype(np.uint64)
            ok = (got["ref_header"] == exp["ref_header"] and got["n_candidates"] == exp["n_candidates"]
                  and np.array_equal(cat, exp["offsets"]))
            res[f"parity_compat{int(compat)}"] = bool(ok)
            assert ok, "sharded index differs from the oracle"
    res["parity"] = bool(res.get("parity_compat1") and res.get("parity_compat0")) if rank == 0 else None
    if args.quick:
        if rank == 0:
            print(json.dumps(res))
        dist.barrier()
        dist.destroy_process_group()
        sys.exit(0)
    # ---- timing: --gib per rank of ONE logical stream of world * gib GiB
    per = args.gib << 30
    block = synth.mp3_like(0xC5, (1 << 28) // 418 - 4)
    block = np.concatenate([block, np.zeros((1 << 28) - block.size, np.uint8)])
    total = per * world
    halo = bd.MPEG_HALO if rank + 1 < world else 0
    dbig = ctx.alloc(per + 256)
    h = ctx.pinned(block.size)
    h.u8[:] = block
    for k in range(per // block.size):
        ctx.lib.blast_memcpy_h2d(ctx.h, dbig.ptr + k * block.size, h.ptr, block.size)
    ctx.lib.blast_memcpy_h2d(ctx.h, dbig.ptr + per, h.ptr, 256)
    ctx.sync()
    import ctypes as C
    from audio_decoder_b200 import _lib
    cap = per // 128 + 4096
    d_pos, d_hdr = ctx.alloc(8 * cap), ctx.alloc(4 * cap)          # preallocated: only the scan is timed
    times = []
    for it in range(5):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        agg_c = _lib.MpegShardAgg()
        assert ctx.lib.blast_mpeg_shard_walk_dev(ctx.h, dbig.ptr, per, halo, C.byref(agg_c)) == 0
        agg = ([int(x) for x in agg_c.exit_state], [int(x) for x in agg_c.count])
        aggs = bd.mpeg_exchange_aggs(agg, None, torch.device("cuda", local))
        folded, tot = bd.mpeg_fold_aggs(aggs)
        entry, _ = folded[rank]
        n = C.c_uint64()
        assert ctx.lib.blast_mpeg_shard_emit_dev(ctx.h, dbig.ptr, per, halo, entry, rank * per, d_pos.ptr, d_hdr.ptr, cap,
                                                 C.byref(n)) == 0
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if it:
            times.append(float(ms))
    if rank == 0:
        ms = float(np.median(times))
        res["scan"] = {"GiB_total": total >> 30, "ms": round(ms, 3), "GBps_scanned": round(total / ms / 1e6, 1), "candidates": tot}
        print(json.dumps(res))
    dist.destroy_process_group()
