"""The CUDA path (through the C ABI) against the committed golden vectors (tests/golden/*.npz)."""
import numpy as np
import pytest

import audio_decoder_b200 as blast
from audio_decoder_b200 import audio_processing as ap, blast_rand as br, file_parsing as fp
from test_golden import DECODE_CASES, MPEG_CASES, RENDER_CASES, SEEDS, conductor_tracks, load, replay_conductor, scene_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = blast.Context(0)
    yield c
    c.close()


def test_decode_golden_gpu(ctx):
    z = load("decode")
    names = DECODE_CASES
    images = [z[n + "_image"] for n in names]
    descs = [fp.probe("wav" if n.startswith("wav") else "aiff", im) for n, im in zip(names, images)]
    outs, _ = fp.decode_batch(ctx, images, descs)
    for n, d, o in zip(names, descs, outs):
        assert np.array_equal(o, z[n + "_samples"]), n
        assert [d.sample_rate, d.num_channels, d.bits_per_sample, d.data_off, d.data_len] == list(z[n + "_meta"]), n


def test_render_golden_gpu(ctx):
    z = load("render")
    for name in RENDER_CASES:
        oc, frames, voices = scene_of(z, name)
        tracks = [ap.Track.from_host(ctx, v["samples"], v["channels"]) for v in voices]
        vp = [ap.VoiceParams(i, True, v["position"], v["velocity"], v["gain"]) for i, v in enumerate(voices)]
        bus, after = ap.render(ctx, tracks, vp, oc, frames)
        assert np.array_equal(bus, z[name + "_bus"]), name
        pos = np.array([a.position for a in after], dtype=np.float32)
        assert np.array_equal(pos.view(np.uint32), z[name + "_final_pos"].view(np.uint32)), name
        # the same scene through the Conductor (Load / set_voice / Start-less activation)
        c = ap.Conductor(ctx, oc, 44100, tracks)
        for i, v in enumerate(voices):
            c.load(i)
            c.set_voice(i, position=v["position"], velocity=v["velocity"], gain=v["gain"], active=True)
        assert np.array_equal(c.coordinate(frames), z[name + "_bus"]), name
        c.close()


def test_rng_golden_gpu(ctx):
    z = load("rng")
    for seed in SEEDS:
        assert list(br.seed_state(seed)) == list(z[f"seed{seed:x}_state"])
        raw, ranged, _ = br.fill(ctx, seed, 0, 1, 64, 0, 100)
        assert np.array_equal(raw[0], z[f"seed{seed:x}_u64"])
        assert np.array_equal(ranged[0], z[f"seed{seed:x}_range_0_100"])
        _, ranged, _ = br.fill(ctx, seed, 0, 1, 64, 50, -7)
        assert np.array_equal(ranged[0], z[f"seed{seed:x}_range_50_m7"])
    raw, _, _ = br.fill(ctx, 42, 1000, 8, 16, 0, 100)
    assert np.array_equal(raw, z["jump_seed42_stride1000_first16"])


def test_mpeg_golden_gpu(ctx):
    z = load("mpeg")
    for name in MPEG_CASES:
        b = z[name + "_bytes"]
        d = ctx.to_device(b)
        pos, hdr = fp.mpeg.scan_dev(ctx, d.ptr, b.size, cap=b.size // 4 + 16)
        assert np.array_equal(pos, z[name + "_pos"]) and np.array_equal(hdr, z[name + "_hdr"]), name
    for h, ok, payload, skip, bitrate in z["header_table"]:
        o = fp.mpeg_header_info(int(h))
        assert bool(o.ok) == bool(ok), hex(h)
        if ok:
            assert o.bitrate == bitrate and o.skip == skip
            assert (int(o.payload_len) if o.frame_len_ok else -1) == payload, hex(h)


def test_conductor_golden_gpu(ctx):
    z = load("conductor")
    tracks = [ap.Track.from_host(ctx, s, ch, 48000) for s, ch in conductor_tracks(z)]
    c = ap.Conductor(ctx, 2, 48000, tracks)
    bus = replay_conductor(c, ap, br.seed_state)
    assert np.array_equal(bus, z["bus"])
    c.close()
