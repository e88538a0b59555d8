"""Python model of K3 `voice_position_scan` (audio_decoder_b200/csrc/render.cu): builds the
arithmetic-segment list of a voice's f32 position trajectory.  Mirrors the CUDA code line by line so
that the ALGORITHM can be validated on the CPU against the sequential recurrence (the CUDA
transliteration itself is validated by the GPU parity tests)."""
import numpy as np

f32 = np.float32
MAXSEG = 160


def bits(x) -> int:
    return int(np.array(x, dtype=np.float32).view(np.uint32))


def f2u_sat(x) -> int:
    x = float(x)
    if x != x or x <= 0:
        return 0
    if x >= 4294967295.0:
        return 0xFFFFFFFF
    return int(x)


def decompose(x):
    b = bits(x)
    sign, E = b >> 31, (b >> 23) & 0xFF
    if E == 255:
        return None
    M = b & 0x7FFFFF
    return sign, E, (M | 0x800000) if E else M


def ulp(E) -> np.float32:
    return f32(2.0 ** ((E if E else 1) - 150))


def seg_eval(p0, d, scale, k):
    # exact in float64: |k*d| < 2^24, scale a power of two, result representable in f32
    if d == 0:
        return f32(p0)
    return f32(np.float64(k * d) * np.float64(scale) + np.float64(p0))


def build_segments(pos, vel, end, total):
    """-> (segments [(step0, p0, d, scale)], final position)"""
    segs = []
    p, vel = f32(pos), f32(vel)
    s = 0
    if total == 0:
        segs.append((0, p, 0, f32(0)))
    with np.errstate(all="ignore"):
        while s < total:
            if f2u_sat(p) >= end:
                segs.append((s, p, 0, f32(0)))
                break
            # integer position, integer velocity: every add below 2^24 is exact whatever binades it crosses, so the run up
            # to the first frozen step (or 2^24, or the end of the epoch) is ONE segment in units of 1.0
            if 1.0 <= float(vel) < 16777216.0 and 0.0 <= float(p) < 16777216.0:
                pi, vv = int(float(p)), int(float(vel))
                if float(pi) == float(p) and float(vv) == float(vel):
                    kmax = (16777215 - pi) // vv
                    kf = (end - pi + vv - 1) // vv                       # first frozen step (end > pi here)
                    kmax = min(kmax, kf, total - s)
                    if kmax >= 1:
                        segs.append((s, p, vv, f32(1.0)))
                        p = f32(pi + kmax * vv)
                        s += kmax
                        continue
            p1 = f32(p + vel)
            if bits(p1) == bits(p):
                segs.append((s, p, 0, f32(0)))
                break
            p2 = f32(p1 + vel)
            a, b, c = decompose(p), decompose(p1), decompose(p2)
            run = a and b and c and a[0] == b[0] == c[0] and a[1] == b[1] == c[1]
            if run:
                s0, E0, q0 = a
                q1, q2 = b[2], c[2]
                d = q2 - q1
                from_p = (q1 - q0) == d
                qs = q0 if from_p else q1
                s_run = s if from_p else s + 1
                p_run = p if from_p else p1
                if d == 0:
                    kmax = 0xFFFFFFFF
                elif d > 0:
                    qhi = (0xFFFFFF if E0 else 0x7FFFFF) - 1
                    kmax = (qhi - qs) // d if qs <= qhi else 0
                else:
                    qlo = 0x800001 if E0 else 0
                    kmax = (qs - qlo) // (-d) if qs >= qlo else 0
                if s0 == 0 and d > 0:
                    e = (E0 if E0 else 1) - 150
                    if e >= 0:
                        thr = 1 if e >= 32 else (end + (1 << e) - 1) >> e
                    else:
                        thr = (1 << 64) - 1 if -e >= 40 else end << (-e)
                    if thr > qs:
                        kf = (thr - qs + d - 1) // d
                        kmax = min(kmax, kf)
                    else:
                        kmax = 0
                kmax = min(kmax, total - s_run)
                if kmax >= 1:
                    if not from_p:
                        segs.append((s, p, 0, f32(0)))
                    ds = -d if s0 else d
                    sc = ulp(E0)
                    segs.append((s_run, p_run, ds, sc))
                    p = seg_eval(p_run, ds, sc, kmax)
                    s = s_run + kmax
                    continue
            segs.append((s, p, 0, f32(0)))
            p = p1
            s += 1
    return segs, p


def expand(segs, total):
    """positions for steps 0..total-1 from the segment list (vectorised)"""
    out = np.empty(total, dtype=np.float32)
    for i, (step0, p0, d, scale) in enumerate(segs):
        nxt = segs[i + 1][0] if i + 1 < len(segs) else total
        nxt = min(nxt, total)
        if nxt <= step0:
            continue
        k = np.arange(nxt - step0, dtype=np.float64)
        if d == 0:
            out[step0:nxt] = p0
        else:
            out[step0:nxt] = (k * d * np.float64(scale) + np.float64(p0)).astype(np.float32)
    return out


def warp_search(lo, hi, pred):
    """Model of K3a's 32-way search (seq_event_scan): first c in [lo, hi) with pred(c), pred monotone (False..True);
    hi if there is none.  Each round evaluates 32 probes (the lanes of the voice's warp) and keeps the sub-range
    between the last False and the first True probe."""
    rounds = 0
    while hi - lo > 32:
        step = (hi - lo) // 32
        b = [pred(lo + lane * step) for lane in range(32)]
        rounds += 1
        if not any(b):
            lo = lo + 31 * step + 1
        else:
            f = b.index(True)
            hi = lo + f * step
            if f == 0:
                break
            lo = lo + (f - 1) * step + 1
    if hi > lo:
        b = [(lo + lane < hi) and pred(lo + lane) for lane in range(32)]
        rounds += 1
        lo = lo + b.index(True) if any(b) else hi
    return lo, rounds


def epochs_put_capped(tpl_step0, epochs, n0=0, last0=None):
    """K3's template epochs as the original per-segment put_capped() sequence: -> list of step0 per segment index"""
    out = [None] * n0
    n = n0
    for base, n_adv in epochs:
        t = 0
        while t < len(tpl_step0) and (t == 0 or tpl_step0[t] < n_adv):
            step0 = base + tpl_step0[t]
            if step0 == last0 and n > 0:
                out[n - 1] = (step0, t)
            else:
                last0 = step0
                out.append((step0, t))
                n += 1
            t += 1
    return out


def epochs_booked(tpl_step0, epochs, n0=0, last0=None):
    """the same as booked by K3's walking lane (first index, count, step offset) and copied by the warp afterwards"""
    cmds, n = [], n0
    for base, n_adv in epochs:
        lo, hi = 1, len(tpl_step0)
        while lo < hi:
            mid = (lo + hi) >> 1
            if tpl_step0[mid] < n_adv:
                lo = mid + 1
            else:
                hi = mid
        start = n - 1 if (base == last0 and n > 0) else n
        cmds.append((start, lo, base))
        n = start + lo
        last0 = base + tpl_step0[lo - 1]
    out = [None] * n
    for start, count, base in cmds:
        for t in range(count):
            out[start + t] = (base + tpl_step0[t], t)
    return out
