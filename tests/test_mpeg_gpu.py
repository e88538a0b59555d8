"""GPU parity: MPEG sync scan (K7), header vote / frame index (K8) and payload gather through the C ABI
vs the oracle's literal mpeg::parse.  Bit-exact."""
import numpy as np
import pytest

import audio_decoder_b200 as blast
from audio_decoder_b200 import file_parsing as fp
import oracle
import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = blast.Context(0)
    yield c
    c.close()


def gpu_scan(ctx, buf):
    a = np.ascontiguousarray(buf, dtype=np.uint8)
    d = ctx.to_device(a) if a.size else ctx.alloc(16)
    return fp.mpeg.scan_dev(ctx, d.ptr, a.size, cap=a.size // 4 + 16)


def same_scan(ctx, buf):
    pos, hdr = gpu_scan(ctx, buf)
    epos, ehdr = oracle.mpeg_sync_scan(buf)
    assert np.array_equal(pos, epos), (len(pos), len(epos), int(np.argmax(pos[:min(len(pos), len(epos))] != epos[:min(len(pos), len(epos))])) if len(pos) and len(epos) else -1)
    assert np.array_equal(hdr, ehdr)
    return pos, hdr


def test_scan_examples(ctx):
    pos, hdr = same_scan(ctx, np.array([0xFF, 0xFF, 0xFB, 0x90, 0x64, 0, 0, 0, 0, 0], dtype=np.uint8))
    assert list(pos) == [0] and list(hdr) == [0xFFFFFB90]
    pos, hdr = same_scan(ctx, np.array([0xFF] * 10 + [0] * 6, dtype=np.uint8))
    assert list(pos) == [0, 4, 8]
    same_scan(ctx, np.array([0, 0xFF, 0xE0, 0], dtype=np.uint8))          # truncated trailing candidate dropped
    same_scan(ctx, np.zeros(1, dtype=np.uint8))
    with pytest.raises(blast.ReferencePanic):
        gpu_scan(ctx, np.array([0, 0, 0xFF], dtype=np.uint8))
    # a trailing 0xFF that is skipped as part of a header does not panic
    b = np.array([0, 0xFF, 0xE0, 1, 0xFF], dtype=np.uint8)
    same_scan(ctx, b)


@pytest.mark.parametrize("alphabet", ["dense", "random", "ffrun"])
def test_scan_random_buffers(ctx, alphabet):
    rng = np.random.default_rng({"dense": 1, "random": 2, "ffrun": 3}[alphabet])
    import os
    for trial in range(int(os.environ.get("BLAST_FUZZ_TRIALS", "40"))):
        n = int(rng.choice([1, 2, 3, 4, 5, 63, 64, 65, 127, 128, 129, 2047, 2048, 2049, 16383, 16384, 16385, 16387,
                            32767, 32768, 32769, 32768 + 61, 65535, 65536, 65537, 70001, 300007]))
        if alphabet == "dense":
            b = rng.choice(np.array([0xFF, 0xFF, 0xE0, 0xFB, 0x00, 0x90, 0xF3], dtype=np.uint8), size=n)
        elif alphabet == "random":
            b = rng.integers(0, 256, size=n, dtype=np.uint8)
        else:
            b = np.full(n, 0xFF, dtype=np.uint8)
            holes = rng.integers(0, n, size=max(1, n // 50))
            b[holes] = rng.integers(0, 256, size=len(holes), dtype=np.uint8)
        b[-1] = 0
        same_scan(ctx, b)


def test_scan_boundaries_every_offset(ctx):
    """a lone header at every offset around the 64-byte chunk, warp-round (2 KiB) and tile (32 KiB) boundaries,
    with and without a second sync 1..4 bytes later (the greedy skip)"""
    for centre in (64, 2048, 16384, 32768, 65536):
        for off in range(centre - 6, centre + 6):
            for gap in (0, 1, 2, 3, 4):
                b = np.zeros(centre + 64, dtype=np.uint8)
                b[off:off + 4] = [0xFF, 0xFB, 0x90, 0x64]
                if gap:
                    b[off + gap:off + gap + 2] = [0xFF, 0xE3]
                same_scan(ctx, b)


def test_scan_long_ff_run_crosses_many_tiles(ctx):
    """adversarial: 200 KiB of 0xFF (one giant cluster: every 4th byte from its start is taken), started at
    each phase relative to the tile grid"""
    for lead in (0, 1, 2, 3, 5):
        b = np.concatenate([np.zeros(lead, np.uint8), np.full(200 * 1024 + 3, 0xFF, np.uint8), np.zeros(8, np.uint8)])
        pos, _ = same_scan(ctx, b)
        assert np.array_equal(pos[:1000], lead + 4 * np.arange(1000, dtype=np.uint64))


def test_index_and_payload_vs_oracle(ctx):
    for seed, n_frames in [(0xC5, 300), (7, 2000), (8, 40000)]:
        buf = synth.mp3_like(seed, n_frames)
        for compat in (True, False):
            exp = oracle.mpeg_parse(buf, reference_compat=compat)
            got = fp.mpeg.parse_bytes(buf, ctx, reference_compat=compat)
            assert got["ref_header"] == exp["ref_header"] == 0xFFFB9064
            assert got["n_candidates"] == exp["n_candidates"]
            assert np.array_equal(got["offsets"], exp["offsets"]), (seed, compat)
            assert np.array_equal(got["payload"], exp["payload"]), (seed, compat)
        # the duplicate-first quirk: exactly one extra entry per distinct valid header value
        a = fp.mpeg.parse_bytes(buf, ctx, reference_compat=True, want_payload=False)["offsets"]
        b = fp.mpeg.parse_bytes(buf, ctx, reference_compat=False, want_payload=False)["offsets"]
        assert len(a) > len(b) and np.all(np.diff(b.astype(np.int64)) > 0) and np.all(np.diff(a.astype(np.int64)) >= 0)


def test_index_reference_panics(ctx):
    with pytest.raises(blast.ReferencePanic):                     # no candidates at all
        fp.mpeg.parse_bytes(np.zeros(1000, np.uint8), ctx)
    junk = np.zeros(1000, np.uint8)
    junk[10:14] = [0xFF, 0xE3, 0xA0, 0x00]                        # only an unparsable header (version bits 01)
    with pytest.raises(blast.ReferencePanic):
        fp.mpeg.parse_bytes(junk, ctx)
    short = np.zeros(100, np.uint8)
    short[90:94] = [0xFF, 0xFB, 0x90, 0x64]                       # payload (257 B) runs past EOF
    with pytest.raises(blast.ReferencePanic):
        fp.mpeg.parse_bytes(short, ctx)
    with pytest.raises(oracle.OracleError):
        oracle.mpeg_parse(short)
    got = fp.mpeg.parse_bytes(short, ctx, reference_compat=False, want_payload=False)
    assert list(got["offsets"]) == [90]


def test_c5_scaled_properties(ctx):
    """C5-shaped stream scaled to 96 MiB (the oracle scans it in a second): full candidate list and frame index
    equal the oracle's; sortedness; idempotence of the device-resident path"""
    n_frames = (96 << 20) // 418
    buf = synth.mp3_like(0xC5, n_frames)
    d = ctx.to_device(buf)
    pos, hdr = fp.mpeg.scan_dev(ctx, d.ptr, buf.size)
    epos, ehdr = oracle.mpeg_sync_scan(buf)
    assert np.array_equal(pos, epos) and np.array_equal(hdr, ehdr)
    idx = fp.mpeg.index_dev(ctx, d.ptr, buf.size, reference_compat=True)
    exp = oracle.mpeg_parse(buf, reference_compat=True, want_payload=False)
    assert idx["ref_header"] == exp["ref_header"] and idx["n_candidates"] == exp["n_candidates"] == len(epos)
    assert np.array_equal(idx["offsets"], exp["offsets"])
    assert np.all(np.diff(idx["offsets"].astype(np.int64)) >= 0)
    again = fp.mpeg.index_dev(ctx, d.ptr, buf.size, reference_compat=True)
    assert np.array_equal(again["offsets"], idx["offsets"])


def _sharded_scan_one_gpu(ctx, b, world):
    """every range gets its own device buffer (range + 16 halo bytes), as on `world` GPUs; the two phases run per range"""
    from audio_decoder_b200 import distributed as bd
    ranges = bd.mpeg_plan_ranges(b.size, world)
    bufs = [ctx.to_device(b[a:a + own + halo]) if own else None for a, own, halo in ranges]
    aggs = [fp.mpeg.shard_walk_dev(ctx, d.ptr, own, halo) if own else ([0, 1, 2, 3], [0, 0, 0, 0])
            for d, (a, own, halo) in zip(bufs, ranges)]
    folded, total = bd.mpeg_fold_aggs(aggs)
    pos, hdr = [], []
    for d, (a, own, halo), agg, (entry, before) in zip(bufs, ranges, aggs, folded):
        if not own:
            continue
        assert fp.mpeg.shard_walk_dev(ctx, d.ptr, own, halo) == agg          # phase 1 again: the scratch is shared here
        d_pos, d_hdr = fp.mpeg.shard_emit_dev(ctx, d.ptr, own, halo, entry, a, agg[1][entry])
        pos.append(d_pos.download(np.uint64, agg[1][entry]))
        hdr.append(d_hdr.download(np.uint32, agg[1][entry]))
    return np.concatenate(pos), np.concatenate(hdr), total


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_scan_equals_whole_scan(ctx, world):
    rng = np.random.default_rng(40 + world)
    cases = [synth.mp3_like(9, 700),                                                       # ~290 KB of frames
             np.concatenate([np.zeros(2, np.uint8), np.full(5 * 32768 + 11, 0xFF, np.uint8), np.zeros(9, np.uint8)]),   # flood
             np.concatenate([rng.choice(np.array([0xFF, 0xFF, 0xE0, 0xFB, 0x00, 0x90], dtype=np.uint8), size=4 * 32768 + 5000),
                             np.zeros(4, np.uint8)])]
    # a sync straddling every range boundary by 1..3 bytes
    edge = np.zeros(9 * 32768, dtype=np.uint8)
    for k in range(1, 9):
        off = k * 32768 - (k % 4)
        edge[off:off + 4] = [0xFF, 0xFB, 0x90, 0x64]
    cases.append(edge)
    for b in cases:
        epos, ehdr = oracle.mpeg_sync_scan(b)
        pos, hdr, total = _sharded_scan_one_gpu(ctx, b, world)
        assert total == len(epos)
        assert np.array_equal(pos, epos) and np.array_equal(hdr, ehdr)


def test_sharded_index_world1_equals_index(ctx):
    from audio_decoder_b200 import distributed as bd
    buf = synth.mp3_like(0xC5, 30000)
    d = ctx.to_device(buf)
    for compat in (True, False):
        exp = oracle.mpeg_parse(buf, reference_compat=compat, want_payload=False)
        got = bd.ShardedMpegIndex(ctx, 0, 1).run(d.ptr, buf.size, reference_compat=compat)
        assert got["ref_header"] == exp["ref_header"] and got["n_candidates"] == exp["n_candidates"]
        assert np.array_equal(got["d_offsets"].download(np.uint64, got["n_offsets"]), exp["offsets"])


def test_offsets_beyond_4GiB(ctx):
    """C5's point is 64-bit positions (mpeg.rs:17-50 indexes with usize): a 5 GiB + stream of zeros with frame clusters
    at the start, straddling 2^32 and at the very end.  Zeros hold no sync byte and leave the scan in its idle state, so
    the candidates / frame offsets are those of the three clusters laid end to end (each followed by zeros) — which the
    oracle parses — translated to where the clusters lie in the large stream."""
    total = (5 << 30) + 123_457
    rng = np.random.default_rng(0x5C5)
    wins = []                                   # (start in the large stream, bytes)
    for k in range(3):
        w = synth.mp3_like(0xC50 + k, (2 << 20) // 418, tail=0)
        start = [0, (1 << 32) - (1 << 20) - 77, total - 1100 - w.size][k]
        w[int(rng.integers(1000, 5000))] = 0xFF            # some stray sync bytes
        if k == 1:
            w[100_000:100_300] = 0xFF                      # a 0xFF run in the cluster that straddles 2^32
            w[100_300] = 0
        wins.append((start, w))
    assert wins[1][0] < (1 << 32) < wins[1][0] + wins[1][1].size and wins[2][0] + wins[2][1].size + 1100 == total
    gap = 4096                                             # zeros between the clusters of the small stream (> any payload)
    small, base = [], []
    at = 0
    for start, w in wins:
        base.append((at, start, w.size))
        small += [w, np.zeros(gap if start != wins[-1][0] else 1100, np.uint8)]
        at += w.size + gap
    small = np.concatenate(small)

    def translate(pos):
        out = np.empty_like(pos)
        for s_at, b_at, n in base:
            m = (pos >= s_at) & (pos < s_at + n + gap)
            out[m] = pos[m] - s_at + b_at
        return out

    d = ctx.alloc(total + 256)
    d.zero()
    for start, w in wins:
        d.upload(w, offset=start)
    epos, ehdr = oracle.mpeg_sync_scan(small)
    pos, hdr = fp.mpeg.scan_dev(ctx, d.ptr, total, cap=len(epos) + 1024)
    assert np.array_equal(pos, translate(epos)) and np.array_equal(hdr, ehdr)
    assert int(np.count_nonzero(pos >= (1 << 32))) > 5000 and pos.max() > (5 << 30)
    for compat in (True, False):
        exp = oracle.mpeg_parse(small, reference_compat=compat, want_payload=False)
        got = fp.mpeg.index_dev(ctx, d.ptr, total, reference_compat=compat)
        assert got["ref_header"] == exp["ref_header"] and got["n_candidates"] == exp["n_candidates"]
        assert np.array_equal(got["offsets"], translate(exp["offsets"])), compat
    d.free()
    ctx.trim()
