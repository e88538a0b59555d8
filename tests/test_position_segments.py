"""The time-parallel position formulation (K3) against the sequential f32 recurrence of the reference
(engine.rs:446 `position += velocity`), on the CPU: random velocities incl. rounding ties, tiny and huge
ratios, negative velocities, zero crossings, the 2^24 stall, freezing at `end`, NaN / inf."""
import numpy as np
import pytest

import oracle
import seg_model as sm

f32 = np.float32


def _check(p0, v, end, total):
    segs, final = sm.build_segments(p0, v, end, total)
    assert len(segs) <= sm.MAXSEG, (p0, v, len(segs))
    ref = oracle.position_walk(p0, v, end, total)
    got = sm.expand(segs, total)
    a, b = got.view(np.uint32), ref[:total].view(np.uint32)
    # -0.0 vs +0.0 is the one tolerated difference (render output cannot tell them apart)
    bad = (a != b) & ~((got == 0) & (ref[:total] == 0))
    assert not bad.any(), (p0, v, end, int(np.argmax(bad)), got[np.argmax(bad)], ref[np.argmax(bad)])
    fb, rb = sm.bits(final), int(ref[total:total + 1].view(np.uint32)[0])
    assert fb == rb or (float(final) == 0.0 and float(ref[total]) == 0.0), (p0, v, end, final, ref[total])
    return len(segs)


def test_named_cases():
    N = 1 << 18
    big = 1 << 31
    assert _check(0.0, 1.0, big, N) <= 48
    _check(16777000.0, 1.0, big, 1000)            # runs into the 2^24 stall
    _check(16777215.0, 3.0, big, 1000)            # v/u = 1.5 tie at 2^24
    _check(0.0, 0.75, 1000, N)                    # freezes at end
    _check(0.0, 0.3, 5, 100)
    _check(100.0, -1.0, big, 500)                 # down through zero, negative forever
    _check(100.25, -0.37, big, N)
    _check(0.0, -1.0, big, 100)
    _check(5.0, 0.0, big, 100)
    _check(1e9, 1.0, 1 << 31, 100)                # velocity below half an ulp: fixed point
    _check(7.0, 3.0, 0, 10)                       # end == 0: frozen at step 0
    _check(0.0, 1e-30, big, N)
    _check(0.0, 1e-45, big, 1000)                 # denormal increments are exact
    _check(1e-40, 1e-42, big, 5000)
    _check(0.0, 1e30, big, 100)                   # frozen after one step (idx saturates)
    _check(3.0, float("inf"), big, 10)
    _check(3.0, float("nan"), big, 10)
    _check(float("nan"), 1.0, big, 10)
    _check(float("-inf"), 1.0, big, 10)
    _check(-0.0, 0.0, big, 10)
    _check(2.5, 1.0, big, 1 << 16)                # v == 1 with a fractional start
    _check(0.1, 1.0, big, 1 << 16)


def test_random_velocities_with_ties():
    rng = np.random.default_rng(123)
    N = 200_000
    big = (1 << 31) - 1
    worst = 0
    for i in range(400):
        kind = i % 8
        if kind == 0:
            v = float(f32(rng.uniform(0.5, 1.5)))
        elif kind == 1:
            v = float(f32(rng.integers(1, 1 << 24)) * f32(2.0) ** int(rng.integers(-30, 3)))   # 24-bit mantissas: ties
        elif kind == 2:
            v = float(f32(rng.uniform(0, 1)) * f32(10.0) ** int(rng.integers(-12, 6)))
        elif kind == 3:
            v = -float(f32(rng.uniform(0.01, 3.0)))
        elif kind == 4:
            v = float(f32(2.0) ** int(rng.integers(-10, 4)) * f32(1.5))                         # k + 1/2 patterns
        elif kind == 5:
            v = float(f32(rng.integers(1, 64)) / f32(rng.integers(1, 64)))
        elif kind == 6:
            v = float(f32(rng.normal()))
        else:
            v = float(f32(rng.uniform(0.9999, 1.0001)))
        p0 = float(f32(rng.choice([0.0, 0.0, 0.5, 123.456, 16777000.0, 1e-3, 70000.7, -5.5])))
        end = int(rng.choice([big, big, 100_000, 5000]))
        worst = max(worst, _check(p0, v, end, N))
    assert worst <= sm.MAXSEG


@pytest.mark.parametrize("v", [1.0, 0.5, 1.5, 0.3, 3.0, 0.1, 7.0 / 3.0])
def test_long_runs(v):
    _check(0.0, v, (1 << 31) - 1, 1 << 21)


def test_integer_velocity_runs_are_one_segment():
    """integer velocity from an integer position: one segment in units of 1.0 up to the first frozen step, the 2^24
    limit or the end of the walk — and still the reference's trajectory bit for bit"""
    big = (1 << 31) - 1
    segs, _ = sm.build_segments(0.0, 1.0, big, 1 << 20)
    assert len(segs) == 1 and segs[0][2] == 1 and float(segs[0][3]) == 1.0
    assert _check(0.0, 1.0, big, 1 << 20) == 1
    assert _check(0.0, 1.0, 720_000 - 1, 720_000) <= 2            # C2: the clip ends with the render (frozen tail segment)
    assert _check(7.0, 3.0, big, 1 << 18) == 1
    assert _check(0.0, 2.0, 5000, 1 << 14) == 2                   # freezes at the clip end
    _check(16777000.0, 1.0, big, 1000)                            # leaves the fast path at 2^24 (stall)
    _check(16777210.0, 5.0, big, 100)
    _check(0.0, 65536.0, big, 1000)
    _check(3.0, 16777215.0, big, 10)
    rng = np.random.default_rng(77)
    for _ in range(200):
        v = float(rng.choice([1.0, 2.0, 3.0, 5.0, 17.0, 1000.0, 65536.0, 8388608.0]))
        p0 = float(rng.integers(0, 1 << 24))
        end = int(rng.choice([big, 100_000, 5000, 3]))
        _check(p0, v, end, int(rng.integers(1, 50_000)))


def test_warp_search_equals_bisection():
    """K3a finds the first call whose tick quotient reaches a value by a 32-way search over the warp's lanes"""
    rng = np.random.default_rng(5)
    for _ in range(2000):
        lo = int(rng.integers(0, 1000))
        hi = lo + int(rng.choice([0, 1, 2, 31, 32, 33, 64, 1000, 1 << 21]))
        first = int(rng.integers(lo - 5, hi + 5))                   # may lie outside the range (-> lo, or none -> hi)
        got, rounds = sm.warp_search(lo, hi, lambda c: c >= first)
        assert got == min(max(first, lo), hi), (lo, hi, first, got)
        assert rounds <= 6


def test_template_epoch_bookkeeping_equals_per_segment_puts():
    """K3 with Seq processes: booking (first index, count, offset) per epoch + a parallel copy gives the segment list
    of the per-segment put_capped() walk, zero-step epochs (replaced by their successor) included"""
    rng = np.random.default_rng(11)
    for _ in range(500):
        n_tpl = int(rng.integers(1, 40))
        tpl = [0] + sorted(set(int(x) for x in rng.integers(1, 5000, size=n_tpl - 1)))
        epochs, base = [], int(rng.integers(0, 100))
        for _e in range(int(rng.integers(1, 30))):
            n_adv = int(rng.choice([0, 0, 1, 2, 7, 100, 3000, 6000]))
            epochs.append((base, n_adv))
            base += n_adv
        n0 = int(rng.integers(0, 3))
        last0 = epochs[0][0] if (n0 and rng.random() < 0.5) else None
        a = sm.epochs_put_capped(tpl, epochs, n0, last0)
        b = sm.epochs_booked(tpl, epochs, n0, last0)
        assert a[n0 - 1 if (n0 and last0 is not None) else n0:] == b[n0 - 1 if (n0 and last0 is not None) else n0:], (tpl, epochs, n0, last0)
        assert len(a) == len(b)
