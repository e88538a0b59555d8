"""N > 1 host logic on the CPU: world_size-2 gloo process group.  Sharding covers every item once, and the
int32 partial-bus all-reduce + wrap reproduces the single-process i16 wrapping mix bit-exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audio_decoder_b200 import distributed as bd
import oracle


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _scene(seed, nv):
    rng = np.random.default_rng(seed)
    voices = []
    for _ in range(nv):
        ch = int(rng.choice([1, 2]))
        nfr = int(rng.integers(300, 900))
        voices.append(dict(samples=rng.integers(-32768, 32768, size=nfr * ch).astype(np.int16), channels=ch,
                           velocity=float(np.float32(rng.choice([1.0, 0.5, 1.37]))),
                           gain=float(np.float32(rng.choice([1.0, 1.9, 0.3]))), position=0.0, active=True))
    return voices


def _oracle_bus(voices, frames):
    c = oracle.Conductor(2, 44100, [(v["samples"], v["channels"], 44100) for v in voices])
    for i, v in enumerate(voices):
        c.load(i)
        c.set_voice(i, position=v["position"], velocity=v["velocity"], gain=v["gain"], active=v["active"])
    return c.coordinate(frames)


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        voices = _scene(123, 9)
        frames = 500
        mine = [voices[i] for i in bd.shard(len(voices), rank, world)]
        # every rank mixes its shard (oracle = CPU stand-in for the render kernel); loud gains make the
        # i16 accumulator wrap, which is exactly the case the int32 reduction has to get right
        part = torch.from_numpy(_oracle_bus(mine, frames).astype(np.int32))
        bd.all_reduce_partial_bus(part)
        bus = bd.wrap_i16(part).numpy()
        np.save(os.path.join(out_dir, f"bus{rank}.npy"), bus)
        counts = torch.tensor([len(mine)])
        dist.all_reduce(counts)
        assert int(counts) == len(voices)
    finally:
        dist.destroy_process_group()


def test_sharding_covers_everything_once():
    for n in (0, 1, 7, 1024, 4097):
        for world in (1, 2, 4, 8):
            seen = sorted(i for r in range(world) for i in bd.shard(n, r, world))
            assert seen == list(range(n))
            assert sum(bd.shard_counts(n, world)) == n


def test_wrap_is_a_ring_homomorphism():
    rng = np.random.default_rng(0)
    a = rng.integers(-32768, 32768, size=(64, 1000)).astype(np.int16)
    seq = np.zeros(1000, dtype=np.int16)
    for row in a:
        seq = (seq.astype(np.int32) + row).astype(np.int16)          # i16 wrapping accumulate, voice by voice
    assert np.array_equal(bd.wrap_i16(a.astype(np.int32).sum(axis=0, dtype=np.int32)), seq)


def test_two_rank_gloo_bus_reduction(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    full = _oracle_bus(_scene(123, 9), 500)
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"bus{r}.npy"), full)


# ---------------------------------------------------------------------------------- MPEG scan over byte ranges
def _range_agg_cpu(b: np.ndarray, start: int, end: int):
    """CPU stand-in for blast_mpeg_shard_walk_dev: the greedy scan (mpeg.rs:17-50) restricted to [start, end) under
    each of the 4 entry states; bytes after `end` are real data (look-ahead / header bytes)"""
    n = len(b)
    exit_state, count = [], []
    for s in range(4):
        cur, cnt, last_end = start + s, 0, start + s
        while cur < end:
            if b[cur] == 0xFF and cur + 1 < n and (b[cur + 1] & 0xE0) == 0xE0:
                if cur + 3 >= n:
                    break
                cnt += 1
                cur += 4
                last_end = cur
            else:
                cur += 1
        exit_state.append(max(0, max(last_end, start + s) - end) if end > start else s)
        count.append(cnt)
    return exit_state, count


def test_mpeg_plan_and_fold_match_the_sequential_scan():
    rng = np.random.default_rng(5)
    for trial in range(6):
        n = int(rng.choice([10, 32768, 65536 + 17, 5 * 32768 + 3, 300007]))
        if trial % 2:
            b = rng.choice(np.array([0xFF, 0xFF, 0xE0, 0xFB, 0x00], dtype=np.uint8), size=n)      # dense: states spill over
        else:
            b = rng.integers(0, 256, size=n, dtype=np.uint8)
        b[-1] = 0
        total = len(oracle.mpeg_sync_scan(b)[0])
        for world in (1, 2, 3, 8):
            ranges = bd.mpeg_plan_ranges(n, world)
            assert ranges[0][0] == 0 and sum(r[1] for r in ranges) == n
            for (a, own, halo), nxt in zip(ranges, ranges[1:] + [(n, 0, 0)]):
                assert a + own == nxt[0]
                assert halo == 0 or (own % bd.MPEG_RANGE_ALIGN == 0 and halo == min(16, n - a - own))
            aggs = [_range_agg_cpu(b, a, a + own) for a, own, _ in ranges]
            folded, tot = bd.mpeg_fold_aggs(aggs)
            assert tot == total, (n, world, tot, total)
            # entry states: what the sequential scan carries into each range
            pos = oracle.mpeg_sync_scan(b)[0].astype(np.int64)
            for (a, own, _), (entry, before) in zip(ranges, folded):
                prev = pos[pos < a]
                want = max(0, int(prev[-1]) + 4 - a) if len(prev) and own else (0 if own else entry)
                if own:
                    assert entry == want and before == len(prev), (n, world, a, entry, want)


def _mpeg_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b = np.random.default_rng(77).choice(np.array([0xFF, 0xFF, 0xE3, 0x90, 0x00, 0x11], dtype=np.uint8), size=3 * 32768 + 999)
        b[-1] = 0
        a, own, _ = bd.mpeg_plan_ranges(len(b), world)[rank]
        aggs = bd.mpeg_exchange_aggs(_range_agg_cpu(b, a, a + own))
        folded, tot = bd.mpeg_fold_aggs(aggs)
        np.save(os.path.join(out_dir, f"mpeg{rank}.npy"), np.array([folded[rank][0], folded[rank][1], tot], dtype=np.int64))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_mpeg_range_exchange(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_mpeg_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    b = np.random.default_rng(77).choice(np.array([0xFF, 0xFF, 0xE3, 0x90, 0x00, 0x11], dtype=np.uint8), size=3 * 32768 + 999)
    b[-1] = 0
    pos = oracle.mpeg_sync_scan(b)[0].astype(np.int64)
    r0, r1 = (np.load(tmp_path / f"mpeg{r}.npy") for r in range(2))
    a1 = bd.mpeg_plan_ranges(len(b), 2)[1][0]
    assert list(r0) == [0, 0, len(pos)]
    before = int((pos < a1).sum())
    assert list(r1) == [max(0, int(pos[pos < a1][-1]) + 4 - a1), before, len(pos)]
