#!/usr/bin/env python
"""Run under torchrun (one rank per GPU, windows mapped by CUDA IPC): the render's exchange step over peer memory
(distributed.PeerBus -> blast_peer_bus) — fused into the render kernel, and as the two-kernel reduction — against (a) the
NCCL all-reduce + finalize path and (b) a single-GPU render of the whole scene on rank 0, over several steps with changing
gains (exercises the ready / done / ack step counters); the sharded Conductor against the CPU oracle; then timings of the
exchange alone (skipped with --quick, the pytest -m gpu entry)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import audio_decoder_b200 as blast  # noqa: E402
from audio_decoder_b200 import audio_processing as ap, distributed as bd  # noqa: E402

if __name__ == "__main__":
    quick = "--quick" in sys.argv
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = blast.Context(local, stream=torch.cuda.current_stream().cuda_stream)
    rng = np.random.default_rng(7)                                   # the same scene on every rank
    n_voices, frames = 64, 200_003
    clips = [rng.integers(-32768, 32768, size=(frames * 2 + 8) * 2).astype(np.int16) for _ in range(n_voices)]
    tracks = [ap.Track.from_host(ctx, c, 2) for c in clips]
    n = frames * 2
    t_part = torch.empty(n, dtype=torch.int32, device=f"cuda:{local}")
    d_bus2 = ctx.alloc(2 * n)
    res = {"world": world}
    for mode in ("fused", "two_kernel", "begin_reduce"):
        peer = bd.PeerBus(ctx, n, rank, world, fused=(mode == "fused"))
        ok = True
        for step in range(5):
            vps = [ap.VoiceParams(v, True, 0.0, 1.0 if v % 3 else 0.77, float(np.float32(0.3 + 0.1 * step + 0.01 * v)))
                   for v in range(n_voices)]
            mine = [p if v % world == rank else ap.VoiceParams(v, False) for v, p in enumerate(vps)]
            sc = ap.Scene(ctx, tracks, mine, 2)
            if mode != "begin_reduce":
                peer.render_reduce(sc, frames)
            else:
                peer.begin()
                sc.render_partial_dev(frames, peer.part_ptr)
                peer.reduce(n)
            a = peer.download_bus(n) if rank == 0 else None
            peer.check()
            sc.set_voices(mine)
            sc.render_partial_dev(frames, t_part.data_ptr())
            dist.all_reduce(t_part, op=dist.ReduceOp.SUM)
            ap.finalize_bus(ctx, t_part.data_ptr(), d_bus2.ptr, n)
            ctx.sync()
            sc.close()
            if rank == 0:
                b = d_bus2.download(np.int16, n)
                whole, _ = ap.render(ctx, tracks, vps, 2, frames)
                ok = ok and np.array_equal(a, b) and np.array_equal(a, whole)
                assert ok, f"mode {mode} step {step}: peer-memory bus differs"
        res["parity_" + mode] = bool(ok)
        dist.barrier()
        peer.close()
    # ---- the Command-driven Conductor over the ranks (ShardedConductor) against the CPU oracle on rank 0
    sc = bd.ShardedConductor(ctx, 2, 48000, tracks[:6], 60_000, rank, world)
    if rank == 0:
        import oracle
        oc = oracle.Conductor(2, 48000, [(c, 2, 48000) for c in clips[:6]])
        seed_state = oracle.Rng(11).state
    else:
        oc, seed_state = None, None
    box = [seed_state]
    dist.broadcast_object_list(box, src=0)
    seed_state = box[0]
    ok = True
    for c, mod in ((sc, ap), (oc, None)):
        if c is None:
            continue
        m = ap if mod is ap else __import__("oracle")
        for t in range(6):
            c.load(t, m.tempo_repr(mode=m.TM_VOICE, interval=float(300 + 37 * t)))
            c.seq(t, m.tempo_repr(owned=False, mode=m.TM_VOICE, idx=t), 4, [0.0, 2.0], [100.0, 60.0], seed_state)
            c.velocity(t, [1.0, 0.8, 1.3, 1.0, 0.5, 1.0][t])
            c.start(t)
    for nfr, cmd in ((20_000, None), (1, ("velocity", 2, 0.9)), (33_333, ("stop", 4)), (60_000, None)):
        got = sc.coordinate(nfr)
        if rank == 0:
            ok = ok and np.array_equal(got, oc.coordinate(nfr))
            assert ok, "sharded conductor differs from the oracle"
        if cmd:
            getattr(sc, cmd[0])(*cmd[1:])
            if rank == 0:
                getattr(oc, cmd[0])(*cmd[1:])
    res["parity_sharded_conductor"] = bool(ok)
    dist.barrier()
    sc.close()
    if not quick:
        # ---- timing of the exchange alone (partial buses already rendered): C3-sized bus (2^20 frames x 2)
        n = 1 << 21
        d_b = ctx.alloc(2 * n)
        t_p = torch.zeros(n, dtype=torch.int32, device=f"cuda:{local}")
        for name in ("two_kernel", "nccl"):
            peer2 = bd.PeerBus(ctx, n, rank, world) if name != "nccl" else None
            times = []
            for it in range(12):
                dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if peer2 is not None:
                    peer2.begin()
                    peer2.reduce(n)
                    peer2.wait()
                else:
                    dist.all_reduce(t_p, op=dist.ReduceOp.SUM)
                    ap.finalize_bus(ctx, t_p.data_ptr(), d_b.ptr, n)
                e1.record()
                torch.cuda.synchronize()
                ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                if it >= 2:
                    times.append(float(ms))
            res[name + "_us"] = round(1e3 * float(np.median(times)), 1)
            dist.barrier()
            if peer2 is not None:
                peer2.close()
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()
