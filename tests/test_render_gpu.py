"""GPU parity: voice render / mix-down (K3 + K4 + K5) through the C ABI vs the oracle Conductor."""
import numpy as np
import pytest

import audio_decoder_b200 as blast
from audio_decoder_b200 import audio_processing as ap
import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = blast.Context(0)
    yield c
    c.close()


def oracle_render(voices, out_channels, frames, chunks=None):
    """voices: list of dict(samples, channels, position, velocity, gain, active)"""
    c = oracle.Conductor(out_channels, 44100, [(v["samples"], v["channels"], 44100) for v in voices])
    for i, v in enumerate(voices):
        c.load(i)
        c.set_voice(i, position=v["position"], velocity=v["velocity"], gain=v["gain"], active=v["active"])
    if chunks is None:
        out = c.coordinate(frames)
    else:
        out = np.concatenate([c.coordinate(n) for n in chunks])
    return out, [c.get_voice(i).position for i in range(len(voices))]


def gpu_render(ctx, voices, out_channels, frames):
    tracks = [ap.Track.from_host(ctx, v["samples"], v["channels"]) for v in voices]
    vp = [ap.VoiceParams(i, v["active"], v["position"], v["velocity"], v["gain"]) for i, v in enumerate(voices)]
    bus, after = ap.render(ctx, tracks, vp, out_channels, frames)
    return bus, [a.position for a in after]


def same_pos(a, b):
    a, b = np.float32(a), np.float32(b)
    return a.view(np.uint32) == b.view(np.uint32) or (a == 0 and b == 0) or (np.isnan(a) and np.isnan(b))


def V(samples, channels, velocity=1.0, gain=1.0, position=0.0, active=True):
    return dict(samples=np.asarray(samples, dtype=np.int16), channels=channels, velocity=velocity, gain=gain,
                position=position, active=active)


ST = np.array([[1000 * k + 7, -1000 * k - 13] for k in range(16)], dtype=np.int16).reshape(-1)
MONO = np.array([300 * k - 5 for k in range(16)], dtype=np.int16)
KATS = {
    "A": ([V(ST, 2)], [7, -13, 1007, -1013, 2007, -2013, 3007, -3013, 4007, -4013, 5007, -5013]),
    "B": ([V(ST, 2, 0.75, 0.5)], [3, -6, 378, -381, 753, -756, 1128, -1131, 1503, -1506, 1878, -1881]),
    "C": ([V(MONO, 1)], [-5, 295, 595, 895, 1195, 1495, 1795, 2095, 2395, 2695, 2995, 3295]),
    "D": ([V(MONO, 1, 0.3, 1.7)], [-8, 144, 297, 450, 603, 756, 909, 1062, 1215, 1368, 1521, 1674]),
    "E": ([V(ST, 2), V(ST, 2, 0.75, 0.5), V(MONO, 1), V(MONO, 1, 0.3, 1.7)],
          [-3, 420, 2277, -49, 4558, -518, 6839, -987, 9120, -1456, 11401, -1925]),
    "F": ([V(np.full(32, 30000), 2), V(np.full(32, 30000), 2)], [-5536] * 12),
    "G": ([V(np.full(32, 30000), 2, 1.0, 2.0)], [32767] * 12),
    "H": ([V(ST, 2, -1.0, 1.0, position=15.0)], [0] * 12),
}


@pytest.mark.parametrize("name", list(KATS))
def test_render_kats(ctx, name):
    voices, expect = KATS[name]
    bus, _ = gpu_render(ctx, voices, 2, 6)
    assert list(bus) == expect


def test_stereo_voice_on_mono_bus_never_advances(ctx):
    bus, pos = gpu_render(ctx, [V(ST, 2)], 1, 6)
    assert list(bus) == [7] * 6 and pos[0] == 0.0


def test_cast_semantics_pinned(ctx):
    """(sample*gain) as i16: saturating, truncating toward zero, NaN -> 0 (SURVEY hard part 2)"""
    clip = np.array([1, 1, 1, 1], dtype=np.int16)
    for gain, want in [(float("inf"), 32767), (float("-inf"), -32768), (float("nan"), 0), (32768.5, 32767),
                       (-32768.5, -32768), (-0.9, 0), (0.9, 0), (-1.9, -1), (32767.99, 32767), (-32769.0, -32768)]:
        bus, _ = gpu_render(ctx, [V(clip, 1, 1.0, gain)], 1, 1)
        exp, _ = oracle_render([V(clip, 1, 1.0, gain)], 1, 1)
        assert bus[0] == want == exp[0], gain


def _random_voice(rng, long_clip=False):
    ch = int(rng.choice([1, 1, 2, 2, 2, 3, 4]))
    nfr = int(rng.integers(2, 60)) if not long_clip else int(rng.integers(3000, 20000))
    s = rng.integers(-32768, 32768, size=nfr * ch).astype(np.int16)
    vel = float(np.float32(rng.choice([1.0, 1.0, 0.5, 1.5, 0.3, 2.25, -1.0, 0.0, 3.7, 0.999, 1e-3, float("nan")])))
    gain = float(np.float32(rng.choice([1.0, 0.5, 1.7, 2.0, -0.75, 0.001, 40.0])))
    pos = float(np.float32(rng.choice([0.0, 0.0, 0.0, 0.5, 3.25, 1e9, -2.5, nfr - 2.0])))
    return V(s, ch, vel, gain, pos, active=bool(rng.random() < 0.9))


@pytest.mark.parametrize("out_channels", [1, 2, 3, 4])
def test_random_small_scenes(ctx, out_channels):
    rng = np.random.default_rng(100 + out_channels)
    for trial in range(25):
        voices = [_random_voice(rng) for _ in range(int(rng.integers(1, 9)))]
        frames = int(rng.integers(1, 80))
        bus, pos = gpu_render(ctx, voices, out_channels, frames)
        exp, epos = oracle_render(voices, out_channels, frames)
        assert np.array_equal(bus, exp), (out_channels, trial)
        assert all(same_pos(a, b) for a, b in zip(pos, epos)), (out_channels, trial, pos, epos)


@pytest.mark.parametrize("out_channels", [1, 2])
def test_multi_tile_scenes_and_freeze(ctx, out_channels):
    """several 2048-frame tiles, clips that end mid-render (freeze), segment boundaries inside tiles"""
    rng = np.random.default_rng(7 + out_channels)
    for trial in range(6):
        voices = [_random_voice(rng, long_clip=True) for _ in range(int(rng.integers(2, 24)))]
        frames = int(rng.integers(4097, 9000))
        bus, pos = gpu_render(ctx, voices, out_channels, frames)
        exp, epos = oracle_render(voices, out_channels, frames)
        assert np.array_equal(bus, exp), (out_channels, trial, int(np.argmax(bus != exp)))
        assert all(same_pos(a, b) for a, b in zip(pos, epos))


def test_binade_crossings_and_stall(ctx):
    """positions near 2^24 (the f32 stall) and velocities that tie: exact trajectory required"""
    n = 1 << 15
    clip = (np.arange(n * 2) % 2001 - 1000).astype(np.int16)
    voices = [V(clip, 2, 1.0, 1.0, position=0.0), V(clip, 2, 3.0, 0.5, position=5.0), V(clip, 2, 0.1, 1.0),
              V(clip, 1, 1.0, 1.0, position=100.0), V(clip, 1, 0.37, 0.8), V(clip, 2, 7.0 / 3.0, 0.3),
              V(clip, 2, 1.0, 1.0, position=16777000.0)]
    frames = 10000
    bus, pos = gpu_render(ctx, voices, 2, frames)
    exp, epos = oracle_render(voices, 2, frames)
    assert np.array_equal(bus, exp)
    assert all(same_pos(a, b) for a, b in zip(pos, epos))


@pytest.mark.parametrize("out_channels", [1, 2])
def test_integer_velocity_runs(ctx, out_channels):
    """integer velocity from an integer position: K3 writes ONE segment in units of 1.0 up to the first frozen step
    (or 2^24).  Clips that end inside the render, start positions next to the clip end, large integer velocities, mono
    and stereo voices, several tiles."""
    rng = np.random.default_rng(2024 + out_channels)
    for trial in range(5):
        voices = []
        for _ in range(12):
            ch = int(rng.choice([1, 2]))
            nfr = int(rng.choice([3, 100, 5000, 20000, 70000]))
            vel = float(rng.choice([1.0, 1.0, 2.0, 3.0, 5.0, 17.0, 1000.0, 65536.0]))
            pos = float(rng.choice([0.0, 0.0, 1.0, 7.0, float(nfr - 1), float(nfr), float(max(0, nfr - 4097)), 4096.0]))
            gain = float(np.float32(rng.choice([1.0, 0.5, 0.77])))
            voices.append(V(rng.integers(-32768, 32768, size=nfr * ch).astype(np.int16), ch, vel, gain, pos))
        frames = int(rng.integers(4097, 12000))
        bus, pos = gpu_render(ctx, voices, out_channels, frames)
        exp, epos = oracle_render(voices, out_channels, frames)
        assert np.array_equal(bus, exp), (out_channels, trial, int(np.argmax(bus != exp)))
        assert all(same_pos(a, b) for a, b in zip(pos, epos)), (pos, epos)


def test_chained_renders_equal_one_long_render(ctx):
    rng = np.random.default_rng(5)
    voices = [_random_voice(rng, long_clip=True) for _ in range(12)]
    tracks = [ap.Track.from_host(ctx, v["samples"], v["channels"]) for v in voices]
    vp = [ap.VoiceParams(i, v["active"], v["position"], v["velocity"], v["gain"]) for i, v in enumerate(voices)]
    sc = ap.Scene(ctx, tracks, vp, 2)
    chunks = [128, 1, 2047, 2048, 3000, 5]
    got = np.concatenate([sc.render(n) for n in chunks])
    exp, epos = oracle_render(voices, 2, sum(chunks), chunks=chunks)
    assert np.array_equal(got, exp)
    assert all(same_pos(a.position, b) for a, b in zip(sc.voices(), epos))
    sc.close()


def test_scaled_c3_scene_vs_oracle_and_linearity(ctx):
    """C3-shaped scene scaled to what the oracle finishes in seconds: 256 voices x 32768 frames, plus the
    size-independent property: mix(all) == wrap16(mix(A) + mix(B)) for a voice partition A | B"""
    rng = np.random.default_rng(0xC3)
    nv, frames = 256, 1 << 15
    voices = []
    for v in range(nv):
        vel = 1.0 if v % 2 == 0 else float(np.float32(0.5 + rng.random()))
        nfr = int(np.ceil(max(1.0, vel) * frames * 1.5)) + 2 if v % 5 else frames // 2      # every 5th clip ends early
        s = rng.integers(-32768, 32768, size=nfr * 2).astype(np.int16)
        voices.append(V(s, 2, vel, float(np.float32(rng.random() * 2.0 ** -5)), 0.0))
    tracks = [ap.Track.from_host(ctx, v["samples"], 2) for v in voices]
    vp = [ap.VoiceParams(i, True, 0.0, v["velocity"], v["gain"]) for i, v in enumerate(voices)]
    bus, _ = ap.render(ctx, tracks, vp, 2, frames)
    exp, _ = oracle_render(voices, 2, frames)
    assert np.array_equal(bus, exp)
    a, _ = ap.render(ctx, tracks, [p if i < 100 else ap.VoiceParams(i, False) for i, p in enumerate(vp)], 2, frames)
    b, _ = ap.render(ctx, tracks, [p if i >= 100 else ap.VoiceParams(i, False) for i, p in enumerate(vp)], 2, frames)
    assert np.array_equal((a.astype(np.int32) + b.astype(np.int32)).astype(np.int16), bus)


def test_reference_panics_become_error_codes(ctx):
    t = ap.Track.from_host(ctx, np.zeros(4, np.int16), 2)
    with pytest.raises(blast.ReferencePanic):
        ap.render(ctx, [t], [ap.VoiceParams(3, True)], 2, 4)                    # track index out of bounds
    empty = ap.Track(ctx.alloc(16), 0, 2)
    with pytest.raises(blast.ReferencePanic):
        ap.render(ctx, [empty], [ap.VoiceParams(0, True)], 2, 4)                # usize underflow in Voice::new
    zero_ch = ap.Track(ctx.alloc(16), 4, 0)
    with pytest.raises(blast.ReferencePanic):
        ap.render(ctx, [zero_ch], [ap.VoiceParams(0, True)], 2, 4)              # divide by zero
    with pytest.raises(blast.BlastError):
        ap.render(ctx, [t], [ap.VoiceParams(0, True)], 9, 4)                    # > 8 output channels
    bus, _ = ap.render(ctx, [t], [], 2, 4)                                      # no voices: silence
    assert list(bus) == [0] * 8


def test_full_size_c3_properties(ctx):
    """BASELINE config 3 at full size — 4,096 stereo voices x 2^20 frames, distinct clips generated on the GPU (17 GB),
    mixed velocities — through the size-independent properties of the mix: linearity over a voice partition (what the
    multi-GPU reduction relies on) and chained renders == one long render (position carry across calls)."""
    import ctypes as C
    from audio_decoder_b200 import blast_rand as br
    voices, frames = 4096, 1 << 20
    clip_words = (int(np.ceil(1.5 * frames)) + 2) * 2
    draws = (clip_words + 3) // 4
    slab = ctx.alloc(voices * draws * 8)
    s = br.Streams(ctx, voices, draws, seed=0xC30000)
    s.fill_dev(draws, 0, 100, slab.ptr, None, None)
    ctx.sync()
    prm = br.fill(ctx, 0xC3, 0, 1, 2 * voices, 0, 100, ranged=False, checks=False)[0][0]
    tracks, vps = [], []
    for v in range(voices):
        u = float((int(prm[2 * v]) >> 11) * 2.0 ** -53)
        w = float((int(prm[2 * v + 1]) >> 11) * 2.0 ** -53)
        buf = blast.DevBuf.__new__(blast.DevBuf)
        buf.ctx, buf.ptr, buf.nbytes = ctx, slab.ptr + v * draws * 8, draws * 8
        buf.free = lambda: None
        tracks.append(ap.Track(buf, clip_words, 2))
        vps.append(ap.VoiceParams(v, True, 0.0, 1.0 if v % 2 == 0 else float(np.float32(0.5) + np.float32(w)),
                                  float(np.float32(u) * np.float32(2.0 ** -7))))
    n = frames * 2
    part = ctx.alloc(4 * n)

    def partial(vp, chunks):
        sc = ap.Scene(ctx, tracks, vp, 2)
        out, off = np.empty(n, dtype=np.int32), 0
        for f in chunks:
            sc.render_partial_dev(f, part.ptr)
            sc.check()
            out[off:off + 2 * f] = part.download(np.int32, 2 * f)
            off += 2 * f
        sc.close()
        return out

    whole = partial(vps, [frames])
    assert np.abs(whole).max() > 1000                                         # something audible was mixed
    # linearity over voice shards v mod 4 (the 4-GPU sharding rule): int32 partial buses add up exactly
    shards = sum(partial([p if i % 4 == r else ap.VoiceParams(i, False) for i, p in enumerate(vps)], [frames]).astype(np.int64)
                 for r in range(4))
    assert np.array_equal(shards.astype(np.int32), whole)
    # three chained renders (ragged cut points) == one long render
    assert np.array_equal(partial(vps, [300_001, 2047, frames - 300_001 - 2047]), whole)


import os as _os


@pytest.mark.parametrize("seed", range(int(_os.environ.get("BLAST_FUZZ_SEEDS", "8"))))
def test_random_scenes_with_weird_floats(ctx, seed):
    """fuzz over the f32 corners of Voice::process: NaN / inf / denormal / negative / huge velocities, positions and
    gains, tiny clips, every channel routing; bus and final positions bit-identical to the oracle"""
    r = np.random.default_rng(9000 + seed)
    oc = int(r.choice([1, 2, 2, 3, 4]))
    frames = int(r.choice([1, 7, 2047, 2048, 2049, 5000]))
    vels = [1.0, 1.0, 0.5, 2.0 ** -20, 1e-39, 3.7, 1000.5, -0.5, float("nan"), float("inf"), float("-inf"), 0.0, -0.0, 0.9999999]
    gains = [1.0, 0.0, -1.0, 1e30, -1e30, float("nan"), float("inf"), 2.0 ** -30, 1.6384, 0.5]
    voices = []
    for _ in range(int(r.integers(1, 7))):
        ch = int(r.choice([1, 2, 2, 3]))
        nfr = int(r.choice([1, 2, 3, 100, 2500, 6000]))
        pos = [0.0, 0.999, 5.5, -1.0, -0.0, 1e8, float("nan"), 16777215.0, 16777216.0, float(nfr - 1), float(nfr), float(nfr + 100),
               float(nfr) - 1.5]
        voices.append(V(r.integers(-32768, 32768, size=nfr * ch).astype(np.int16), ch, float(r.choice(vels)), float(r.choice(gains)),
                        float(r.choice(pos)), bool(r.random() < 0.9)))
    with np.errstate(all="ignore"):
        exp, epos = oracle_render(voices, oc, frames)
        got, gpos = gpu_render(ctx, voices, oc, frames)
    assert np.array_equal(got, exp), (seed, oc, frames)
    for a, b in zip(gpos, epos):
        assert same_pos(a, b), (seed, a, b)


def test_mono_unit_velocity_on_stereo_bus(ctx):
    """a mono voice at velocity 1.0 advances on both channels (L = s[i0 + 2f], R = s[i0 + 2f + 1]): even start positions
    take the packed-pair fast path, odd ones the generic one; gains 1.0 (integer mix) and != 1.0; clips that end inside"""
    rng = np.random.default_rng(77)
    long_ = rng.integers(-32768, 32768, size=40000).astype(np.int16)
    short = rng.integers(-32768, 32768, size=9001).astype(np.int16)
    for frames in (5000, 9000):
        voices = [V(long_, 1, 1.0, 1.0, 0.0), V(long_, 1, 1.0, 0.37, 1.0), V(long_, 1, 1.0, 1.0, 7.0), V(long_, 1, 1.0, 2.5, 4096.0),
                  V(short, 1, 1.0, 1.0, 0.0), V(short, 1, 1.0, 0.9, 3.0)]
        exp, epos = oracle_render(voices, 2, frames)
        got, gpos = gpu_render(ctx, voices, 2, frames)
        assert np.array_equal(got, exp), frames
        assert all(same_pos(a, b) for a, b in zip(gpos, epos))


def test_c3_all_4096_voices_16k_frames_vs_oracle(ctx):
    """SURVEY §8(d) C3 at the full voice count, generated as specified (clips from X128P::new(0xC3_0000 + v), gains and
    velocities from X128P::new(0xC3): odd voices interpolate at 0.5..1.5), over the frames the oracle finishes in
    seconds: 4,096 voices x 16,384 frames against Conductor::coordinate (engine.rs:46-81, 386-448)"""
    import synth
    nv, frames = synth.C3_VOICES, 1 << 14
    params = synth.c3_voice_params(nv)
    clips = []
    for v in range(nv):
        n = synth.c3_clip_frames(v, frames) * 2                       # stereo samples
        raw = oracle.Rng(0xC30000 + v).fill_u64((n + 3) // 4)         # successive next_u64 bytes, little-endian
        clips.append(raw.view(np.int16)[:n].copy())
    tracks = [ap.Track.from_host(ctx, c, 2) for c in clips]
    vp = [ap.VoiceParams(v, True, 0.0, params[v][0], params[v][1]) for v in range(nv)]
    bus, after = ap.render(ctx, tracks, vp, 2, frames)
    exp, epos = oracle_render([V(clips[v], 2, params[v][0], params[v][1]) for v in range(nv)], 2, frames)
    assert np.array_equal(bus, exp), int(np.argmax(bus != exp))
    assert all(same_pos(a.position, b) for a, b in zip(after, epos))
    assert int(np.count_nonzero(bus)) > frames                         # not a silent bus
