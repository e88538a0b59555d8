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
