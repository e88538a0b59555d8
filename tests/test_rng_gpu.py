"""GPU parity: X128P streams (K6) through the C ABI vs the oracle's sequential generator. Bit-exact."""
import numpy as np
import pytest

import audio_decoder_b200 as blast
from audio_decoder_b200 import blast_rand as br
import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = blast.Context(0)
    yield c
    c.close()


def test_kats(ctx):
    kats = {0: ([0x509946a41cd733a3, 0x00885667b1934bfa, 0x1061f9ad258fd5d5, 0x3f8be44897a4317c],
                [31, 0, 6, 24, 37, 83, 14, 72, 75, 76, 89, 93]),
            42: ([0xe6c71559e2525f98, 0xc47d57593d0cfb7a, 0x39de93182b828cf8, 0x7f6298c8e5492240],
                 [90, 76, 22, 49, 68, 29, 29, 0, 55, 98, 70, 77])}
    for seed, (first4, ranged) in kats.items():
        r, g, c = br.fill(ctx, seed, 0, 1, 12, 0, 100)
        assert list(r[0, :4]) == first4
        assert list(g[0]) == ranged


@pytest.mark.parametrize("n_streams,draws,stride", [(1, 1, 7), (5, 33, 33), (32, 16, 16), (33, 17, 1000), (100, 1000, 1000),
                                                    (257, 129, 65536), (64, 48, 0), (40, 250, 1)])
def test_streams_match_sequential(ctx, n_streams, draws, stride):
    lo, hi = -17, 1000
    s = br.Streams(ctx, n_streams, stride, seed=1234)
    out = s.fill(draws, lo, hi)
    for k in range(n_streams):
        g = oracle.Rng(1234)
        g.discard(k * stride)
        exp = oracle.Rng(state=g.state).fill_u64(draws)
        assert np.array_equal(out["raw"][k], exp), k
        assert np.array_equal(out["ranged"][k], oracle.Rng(state=g.state).fill_range(lo, hi, draws)), k
        assert tuple(int(x) for x in out["checks"][k]) == oracle.Rng(state=g.state).checksum(lo, hi, draws)
    # states were advanced in place: a second fill continues the sequences
    out2 = s.fill(3, lo, hi, ranged=False, checks=False)
    g = oracle.Rng(1234)
    g.discard((n_streams - 1) * stride + draws)
    assert np.array_equal(out2["raw"][-1], g.fill_u64(3))


def test_range_quirks(ctx):
    for lo, hi in [(100, 0), (5, 5), (-2**62, 2**62), (0, 2**63 - 1), (-2**63, 2**63 - 1)]:
        r, g, _ = br.fill(ctx, 7, 50, 3, 50, lo, hi)
        for k in range(3):
            o = oracle.Rng(7)
            o.discard(50 * k)
            assert np.array_equal(g[k], o.fill_range(lo, hi, 50)), (lo, hi)


def test_c4_full_size_checksums_and_stream_dumps(ctx):
    """BASELINE config 4 at full size: 2^32 draws over 65,536 jump-ahead streams (stride 65,536): per-stream
    checksums of every stream vs one sequential CPU pass; element-wise dumps of streams {0,1,255,4095,65535}."""
    n, draws = 65536, 65536
    s = br.Streams(ctx, n, draws, seed=42)
    st = s.get_states()
    assert tuple(int(x) for x in st[1]) == (0x7642b3b57ffb2a57, 0xc7c3ecc7c44024b3)
    checks = ctx.alloc(32 * n)
    s.fill_dev(draws, 0, 100, None, None, checks.ptr)
    got = checks.download(np.uint64, 4 * n).reshape(n, 4)
    g = oracle.Rng(42)
    picks = {0, 1, 255, 4095, 65535}
    dumps = {}
    for k in range(n):
        if k in picks:
            state = g.state
            dumps[k] = (oracle.Rng(state=state).fill_u64(draws), oracle.Rng(state=state).fill_range(0, 100, draws))
        exp = g.checksum(0, 100, draws)
        assert tuple(int(x) for x in got[k]) == exp, k
    # element-wise dumps through the materialising path (5 streams re-generated from their jump states)
    for k in sorted(picks):
        one = br.Streams(ctx, 1, 0, state=tuple(int(x) for x in st[k]))
        out = one.fill(draws, 0, 100, checks=False)
        assert np.array_equal(out["raw"][0], dumps[k][0]), k
        assert np.array_equal(out["ranged"][0], dumps[k][1]), k
