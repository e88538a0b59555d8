"""CPU models of the integer / bit tricks K4's consumers and its producer's three-segment cut rely on
(audio_decoder_b200/csrc/render.cu: unpack_pair, consume_stereo_lerp<., true>, the lane-local cut of the producer).
The CUDA transliteration itself is covered by the GPU parity tests; these pin the ARITHMETIC claims on the CPU:

* i16 -> f32 without a conversion instruction: ((w & 0xFFFF) ^ 0x4B008000) is the float 2^23 + 32768 + L and
  0x4B400000 + sext(w >> 16) the float 2^23 + 2^22 + R, for every 16-bit L and R;
* the fraction trick of the interpolated path: with sh <= 23 fraction bits (q & mask) | (150 - sh) << 23 is the float
  2^(23 - sh) + fract, so fract and 1 - fract are exact float adds;
* two or three consecutive segments of a trajectory (tests/seg_model.py, the model of K3) brought to their finest common
  unit 2^-sh reproduce, frame by frame, the index and the fraction the reference computes from the f32 position
  (engine.rs:407, 430-433): idx = trunc(position), frac = position.fract().
"""
import numpy as np
import pytest

import seg_model as sm

f32 = np.float32


def _as_f32(bits_u32):
    return np.asarray(bits_u32, dtype=np.uint32).view(np.float32)


def test_unpack_left_channel_all_values():
    lo = np.arange(65536, dtype=np.uint32)
    got = _as_f32(lo ^ np.uint32(0x4B008000)) + f32(-8421376.0)                    # -(2^23 + 32768)
    want = lo.astype(np.uint16).view(np.int16).astype(np.float32)
    assert np.array_equal(got, want)


def test_unpack_right_channel_all_values():
    hi = np.arange(65536, dtype=np.uint32)
    w = hi << np.uint32(16) | np.uint32(0x1234)                                    # any left half
    sext = (w.view(np.int32) >> 16).astype(np.int64)
    bits = ((np.int64(0x4B400000) + sext) & 0xFFFFFFFF).astype(np.uint32)          # what LEA.HI.SX32 computes
    got = _as_f32(bits) + f32(-12582912.0)                                         # -(2^23 + 2^22)
    want = hi.astype(np.uint16).view(np.int16).astype(np.float32)
    assert np.array_equal(got, want)
    # the packed constant of the one FADD2: low word -(2^23 + 32768), high word -(2^23 + 2^22)
    c = np.array([0xCB008000, 0xCB400000], dtype=np.uint32).view(np.float32)
    assert c[0] == f32(-8421376.0) and c[1] == f32(-12582912.0)


@pytest.mark.parametrize("sh", [0, 1, 5, 12, 22, 23])
def test_fraction_trick(sh):
    rng = np.random.default_rng(sh)
    q = rng.integers(0, 1 << 31, size=4096, dtype=np.int64).astype(np.uint32)
    mask = np.uint32((1 << sh) - 1)
    magic = np.uint32((150 - sh) << 23)
    v = _as_f32((q & mask) | magic)
    frac = v + (-_as_f32(magic))
    om = (_as_f32(magic) + f32(1.0)) - v
    want = (q & mask).astype(np.float64) / float(1 << sh)
    assert np.array_equal(frac.astype(np.float64), want)                           # exact, and what f32::fract gives
    assert np.array_equal(om.astype(np.float64), 1.0 - want)


def _pieces_of_tile(segs, f0, nf):
    """the segments that overlap tile frames [f0, f0 + nf) as (fa, fe, p_a, d, scale) — S == 1 (steps are frames)"""
    out = []
    for k, (step0, p0, d, scale) in enumerate(segs):
        nxt = segs[k + 1][0] if k + 1 < len(segs) else 1 << 62
        ls, le = max(step0, f0), min(nxt, f0 + nf)
        if ls < le:
            out.append((ls - f0, le - f0, sm.seg_eval(p0, d, scale, ls - step0), d, scale))
    return out


def _sh_of(p_a, d, scale):
    if d != 0:
        eb = (sm.bits(scale) >> 23) & 0xFF
        return 127 - eb if 104 <= eb <= 127 else None
    E = (sm.bits(p_a) >> 23) & 0xFF
    s = 0 if float(p_a) == 0.0 else max(0, 150 - E)
    return s if s <= 23 else None


@pytest.mark.parametrize("seed", range(12))
def test_three_segments_in_a_common_unit(seed):
    rng = np.random.default_rng(1000 + seed)
    kFT = 2048
    n_checked = 0
    for _ in range(40):
        vel = f32(rng.uniform(0.5, 1.5))
        start = f32(rng.uniform(0.0, 60000.0))
        total = 6 * kFT
        segs, _ = sm.build_segments(start, vel, 1 << 23, total)
        pos = sm.expand(segs, total)                                               # the f32 trajectory, step by step
        for t in range(total // kFT):
            f0 = t * kFT
            pieces = _pieces_of_tile(segs, f0, kFT)
            if not 2 <= len(pieces) <= 3:
                continue
            shs = [_sh_of(p_a, d, scale) for (_, _, p_a, d, scale) in pieces]
            if any(s is None for s in shs):
                continue
            sh_c = max(shs)
            hi = int(max(float(pos[f0 + fe - 1]) for (_, fe, _, _, _) in pieces))
            if (hi + 2) >> (31 - sh_c):
                continue                                                           # (the producer leaves those to the table path)
            for (fa, fe, p_a, d, scale), sh in zip(pieces, shs):
                qa = int(float(p_a) * float(1 << sh))                              # exact: the significand
                assert float(qa) == float(p_a) * float(1 << sh)
                q0 = ((qa - fa * d) << (sh_c - sh)) & 0xFFFFFFFF
                dd = (d << (sh_c - sh)) & 0xFFFFFFFF
                fl = np.arange(fa, fe, dtype=np.int64)
                q = (q0 + fl * dd) & 0xFFFFFFFF
                p = pos[f0 + fa:f0 + fe].astype(np.float64)
                assert np.array_equal(q >> sh_c, np.trunc(p).astype(np.int64))
                frac = (q & ((1 << sh_c) - 1)).astype(np.float64) / float(1 << sh_c)
                assert np.array_equal(frac, p - np.trunc(p))
                n_checked += 1
    assert n_checked > 20


# ---------------------------------------------------------------------------------------------------------------------
# The producer warp of K4: which lanes issue an item in a round, which pair up, which stage each item takes
# (render.cu, issue_round) and which held tiles the voice's own lane may cut (the segment discovery of the lane-local cut).
K_STAGES = 4


def _issue_round_model(have: int, pairable: int, o: int):
    """-> [(stage item index, first lane, second lane or None)] in issue order, as the 32 lanes compute it"""
    items = []
    lanes = {}
    for lane in range(32):
        lt = (1 << lane) - 1
        is_pair = (pairable >> lane) & 1
        second = bool(is_pair and (bin(pairable & lt).count("1") & 1))
        above = pairable & ~lt & ~(1 << lane) & 0xFFFFFFFF
        partner = ((above & -above).bit_length() - 1) if (is_pair and not second and above) else -1
        lanes[lane] = (second, partner)
    second_mask = sum(1 << l for l, (s, _) in lanes.items() if s)
    item_mask = have & ~second_mask
    for lane in range(32):
        if (item_mask >> lane) & 1:
            irank = bin(item_mask & ((1 << lane) - 1)).count("1")
            items.append((o + irank, lane, lanes[lane][1] if lanes[lane][1] >= 0 else None))
    return items


def test_issue_round_pairs_and_stages():
    rng = np.random.default_rng(77)
    for _ in range(400):
        have = int(rng.integers(0, 1 << 32))
        pairable = int(rng.integers(0, 1 << 32)) & have
        o = int(rng.integers(0, 1000))
        got = _issue_round_model(have, pairable, o)
        # the sequential rule it replaces: lanes in ascending order; a pairable lane takes the next pairable lane along
        want, rest, k = [], have, o
        while rest:
            i = (rest & -rest).bit_length() - 1
            rest &= rest - 1
            i2 = None
            if (pairable >> i) & 1:
                cand = rest & pairable
                if cand:
                    i2 = (cand & -cand).bit_length() - 1
                    rest &= ~(1 << i2)
            want.append((k, i, i2))
            k += 1
        assert got == want
        # the lanes of one wave (K_STAGES consecutive items) take distinct stages
        for w0 in range(0, len(got), K_STAGES):
            stages = [idx % K_STAGES for idx, _, _ in got[w0:w0 + K_STAGES]]
            assert len(set(stages)) == len(stages)


def _lane_local_discovery(step0s, seg_j, last_abs):
    """cnt (1..3 segments that overlap the tile from seg_j on) and whether no further segment starts inside the tile"""
    nseg = len(step0s)
    cnt, nxt = 1, [0xFFFFFFFF] * 3
    for t in (1, 2, 3):
        if cnt == t and seg_j + t < nseg:
            s0 = step0s[seg_j + t]
            nxt[t - 1] = s0
            if t < 3 and s0 <= last_abs:
                cnt = t + 1
    nxt_last = nxt[2] if cnt == 3 else (nxt[1] if cnt == 2 else nxt[0])
    return cnt, nxt_last > last_abs


def test_lane_local_segment_discovery():
    rng = np.random.default_rng(5)
    for _ in range(2000):
        nseg = int(rng.integers(1, 9))
        gaps = rng.integers(1, 1500, size=nseg - 1) if nseg > 1 else np.array([], dtype=np.int64)
        step0s = [0] + list(np.cumsum(gaps).astype(int))
        f0 = int(rng.integers(0, max(1, step0s[-1] + 500)))
        nf = 2048
        last_abs = f0 + nf - 1
        seg_j = max(j for j in range(nseg) if step0s[j] <= f0)                 # the segment that holds the tile's first step
        inside = [j for j in range(seg_j, nseg) if j == seg_j or step0s[j] <= last_abs]
        cnt, covered = _lane_local_discovery(step0s, seg_j, last_abs)
        if len(inside) <= 3:
            assert (cnt, covered) == (len(inside), True)
        else:
            assert cnt == 3 and not covered
