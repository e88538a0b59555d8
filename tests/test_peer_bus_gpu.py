"""GPU parity of the render's exchange step (blast_peer_bus, fused into the render kernel) and of the single-process
multi-GPU group (blast_group): always against the CPU oracle or the single-GPU entry points.

A group may name one device several times, so the multi-member tile protocol (ready / done / ack flags, tile ownership,
remote-looking loads and stores) runs on the driver's single-GPU box too; with >= 2 visible GPUs the same tests also run
over real peer memory, and the one-process-per-GPU variant (CUDA IPC, torchrun) is launched from here."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import audio_decoder_b200 as blast
from audio_decoder_b200 import _lib, audio_processing as ap, blast_rand as br, distributed as bd, file_parsing as fp
from audio_decoder_b200.group import Group, GroupConductor
import oracle
import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus():
    import torch
    return torch.cuda.device_count()


def member_sets():
    """device lists for groups: repeated device 0 always; distinct GPUs when the box has them"""
    sets = [[0], [0, 0], [0, 0, 0]]
    n = n_gpus()
    if n >= 2:
        sets.append([0, 1])
    if n >= 4:
        sets.append([0, 1, 2, 3])
    if n >= 8:
        sets.append(list(range(8)))
    return sets


@pytest.fixture(scope="module")
def ctx():
    c = blast.Context(0)
    yield c
    c.close()


def _scene(rng, n_voices, frames, mono_every=0):
    clips, voices = [], []
    for v in range(n_voices):
        ch = 1 if mono_every and v % mono_every == 0 else 2
        vel = 1.0 if v % 3 else float(np.float32(0.4 + 1.2 * rng.random()))
        n = int(frames * max(vel, 1.0) * (2 if ch == 1 else 1)) + 16 if v % 4 else frames // 3 + 5   # some clips end early
        clips.append((rng.integers(-32768, 32768, size=n * ch).astype(np.int16), ch))
        voices.append(ap.VoiceParams(v, True, 0.0, vel, float(np.float32(rng.random() * 0.2))))
    return clips, voices


def _oracle_bus(clips, voices, oc, frames):
    c = oracle.Conductor(oc, 48000, [(s, ch, 48000) for s, ch in clips])
    for i, v in enumerate(voices):
        c.load(v.track)
        c.set_voice(i, position=v.position, velocity=v.velocity, gain=v.gain, active=v.active)
    return c.coordinate(frames)


@pytest.mark.parametrize("fused", [False, True], ids=["two_kernel", "fused"])
@pytest.mark.parametrize("oc,frames", [(2, 1), (2, 2047), (2, 2048), (2, 40_001), (1, 9_000), (3, 5_000)])
def test_world1_render_reduce_is_the_finalized_render(ctx, oc, frames, fused):
    """world == 1: render + finalize, or (fused) the render kernel finalizes its own tiles — same S16 bus as the oracle"""
    rng = np.random.default_rng(frames + oc)
    clips, voices = _scene(rng, 70, frames, mono_every=7)
    tracks = [ap.Track.from_host(ctx, s, ch) for s, ch in clips]
    sc = ap.Scene(ctx, tracks, voices, oc)
    pb = bd.PeerBus(ctx, frames * oc + 13, 0, 1, fused=fused)
    for _ in range(2):                                   # twice: flags are step counters, the voices are rewound
        sc.restore_dev()
        pb.render_reduce(sc, frames)
        got = pb.download_bus(frames * oc)
        sc.check()
        assert np.array_equal(got, _oracle_bus(clips, voices, oc, frames))
    pb.close()
    sc.close()


@pytest.mark.parametrize("fused", [False, True], ids=["two_kernel", "fused"])
@pytest.mark.parametrize("devices", member_sets(), ids=lambda d: "gpus_" + "_".join(map(str, d)))
def test_group_decode_render_against_the_oracle(devices, fused):
    """main.rs:18-89 + Conductor::coordinate over a group: files decoded on member i mod n, voices rendered where their
    track lives, bus reduced inside the render kernel — bit-identical to the oracle's single-threaded result"""
    rng = np.random.default_rng(len(devices) * 101 + sum(devices))
    n_files = 13
    images, kinds = [], []
    for i in range(n_files):
        n = int(rng.integers(30_000, 60_000)) * 4
        if i % 2:
            images.append(synth.wav_image(100 + i, n)); kinds.append("wav")
        else:
            images.append(synth.aiff_image(100 + i, n, bits=16)); kinds.append("aiff")
    descs = [fp.probe(k, im) for k, im in zip(kinds, images)]
    with Group(devices, fused=fused) as g:
        outs, tracks = g.decode_batch(images, descs)
        exp = [(oracle.wav_parse if k == "wav" else oracle.aiff_parse)(im)[1] for k, im in zip(kinds, images)]
        for o, e in zip(outs, exp):
            assert np.array_equal(o, e)
        for frames in (1, 5_000, 29_999):
            voices = [ap.VoiceParams(i % n_files, True, 0.0, 1.0 if i % 3 else float(np.float32(0.5 + rng.random())),
                                     float(np.float32(0.05 + 0.3 * rng.random()))) for i in range(2 * n_files + 3)]
            for rep in range(2):
                bus = g.render(tracks, n_files, voices, 2, frames)
                assert np.array_equal(bus, _oracle_bus([(e, 2) for e in exp], voices, 2, frames)), (devices, frames, rep)
        # more bus channels than the fused kernel handles: the two-kernel reduction
        bus = g.render(tracks, n_files, voices[:9], 3, 4_100)
        assert np.array_equal(bus, _oracle_bus([(e, 2) for e in exp], voices[:9], 3, 4_100))


@pytest.mark.parametrize("devices", member_sets()[1:], ids=lambda d: "gpus_" + "_".join(map(str, d)))
def test_group_conductor_with_seq_against_the_oracle(devices):
    rng = np.random.default_rng(5 + len(devices))
    clips = [rng.integers(-32768, 32768, size=(70_000 + 8) * 2).astype(np.int16) for _ in range(7)]
    seed_state = oracle.Rng(11).state
    with Group(devices) as g:
        # tracks must live where the sharding rule says: upload clip t on member t mod n
        bufs, tracks = [], (_lib.Track * len(clips))()
        for t, s in enumerate(clips):
            b = g.member_context(t % g.n).to_device(s)
            bufs.append(b)
            tracks[t] = _lib.Track(b.ptr, s.size, 2, 48000)
        gc = GroupConductor(g, 2, 48000, tracks, len(clips))
        oc = oracle.Conductor(2, 48000, [(c, 2, 48000) for c in clips])
        for c, m in ((gc, ap), (oc, oracle)):
            for t in range(len(clips)):
                c.load(t, m.tempo_repr(mode=m.TM_VOICE, interval=float(300 + 37 * t)))
                c.seq(t, m.tempo_repr(owned=False, mode=m.TM_VOICE, idx=t), 4, [0.0, 2.0], [100.0, 60.0], seed_state)
                c.velocity(t, [1.0, 0.8, 1.3, 1.0, 0.5, 1.0, 1.7][t])
                c.start(t)
        for frames, cmd in ((20_000, None), (1, ("velocity", 2, 0.9)), (33_333, ("stop", 4)), (9_000, None)):
            assert np.array_equal(gc.coordinate(frames), oc.coordinate(frames)), (devices, frames)
            if cmd:
                getattr(gc, cmd[0])(*cmd[1:])
                getattr(oc, cmd[0])(*cmd[1:])
        gc.close()
        for b in bufs:
            b.free()


@pytest.mark.parametrize("devices", member_sets()[1:], ids=lambda d: "gpus_" + "_".join(map(str, d)))
def test_group_rng_and_mpeg_equal_the_single_gpu_results(ctx, devices):
    with Group(devices) as g:
        raw, ranged, checks = g.x128p_fill(42, 1000, 37, 257, 0, 100)
        for s_ in (0, 1, 2, 17, 36):
            r = oracle.Rng(42)
            r.discard(1000 * s_)
            assert np.array_equal(raw[s_], r.fill_u64(257))
            r = oracle.Rng(42)
            r.discard(1000 * s_)
            assert np.array_equal(ranged[s_], r.fill_range(0, 100, 257))
        raw1, ranged1, checks1 = br.fill(ctx, 42, 1000, 37, 257, 0, 100)
        assert np.array_equal(raw, raw1) and np.array_equal(ranged, ranged1) and np.array_equal(checks, checks1)
        for n_frames, compat in ((3000, True), (3000, False), (700, True)):
            stream = synth.mp3_like(9 + n_frames, n_frames)
            stream[40_000:40_900] = 0xFF                         # a 0xFF flood across a span boundary
            stream[40_900] = 0x00
            got = g.mpeg_index(stream, reference_compat=compat)
            exp = oracle.mpeg_parse(stream, reference_compat=compat, want_payload=False)
            assert got["ref_header"] == exp["ref_header"]
            assert got["n_candidates"] == exp["n_candidates"]
            assert np.array_equal(got["offsets"], exp["offsets"]), (devices, n_frames, compat)


def test_group_c_harness():
    """the same surface from plain C++ (no Python, no torch in the process): tests/checks/group_harness.cpp"""
    exe = os.path.join(ROOT, "tests", "checks", "_group_harness")
    src = os.path.join(ROOT, "tests", "checks", "group_harness.cpp")
    lib_dir = os.path.join(ROOT, "audio_decoder_b200")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                           "-L", lib_dir, "-lblast_cuda", f"-Wl,-rpath,{lib_dir}"])
    for devs in member_sets()[1:]:
        out = subprocess.run([exe] + [str(d) for d in devs], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "group harness OK" in out.stdout


def _torchrun(n, script, *args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + n), os.path.join(ROOT, "tests", "checks", script), *args]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-4000:]
    return json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])


@pytest.mark.skipif(n_gpus() < 2, reason="needs >= 2 GPUs")
def test_one_process_per_gpu_peer_bus_over_ipc():
    """torchrun, CUDA IPC windows: fused render+reduce, the two-kernel reduction and the sharded Conductor against a
    single-GPU render of the whole scene and against the oracle"""
    for n in [k for k in (2, 4, 8) if k <= n_gpus()]:
        res = _torchrun(n, "peer_bus_check.py", "--quick")
        assert res["parity_fused"] and res["parity_two_kernel"] and res["parity_begin_reduce"] and res["parity_sharded_conductor"], res


@pytest.mark.skipif(n_gpus() < 2, reason="needs >= 2 GPUs")
def test_one_process_per_gpu_mpeg_index():
    for n in [k for k in (2, 8) if k <= n_gpus()]:
        res = _torchrun(n, "mpeg_sharded_check.py", "--quick")
        assert res["parity"], res
