"""GPU parity: the Command-driven Conductor (blast_conductor_*: K3a Seq event scan + K3 + K4 + K5) vs the oracle's
restatement of Conductor::{apply, coordinate}, Voice / Group transport, TempoState and Seq
(engine.rs:36-248, 318-384, 477-542; blast_time.rs:58-161; processes.rs:52-99).  Bus: bit-exact."""
import numpy as np
import pytest

import audio_decoder_b200 as blast
from audio_decoder_b200 import _lib, audio_processing as ap
import oracle

pytestmark = pytest.mark.gpu

SR = 48000


@pytest.fixture(scope="module")
def ctx():
    c = blast.Context(0)
    yield c
    c.close()


def make_tracks(seed, spec):
    """spec: list of (frames, channels) -> list of int16 arrays"""
    r = np.random.default_rng(seed)
    return [(r.integers(-20000, 20000, size=n * ch, dtype=np.int16), ch) for n, ch in spec]


class Pair:
    """The same operations on the oracle Conductor and on the GPU Conductor."""

    def __init__(self, ctx, out_channels, tracks, sample_rate=SR):
        self.ctx = ctx
        self.oc = out_channels
        self.o = oracle.Conductor(out_channels, sample_rate, [(s, ch, sample_rate) for s, ch in tracks])
        self.dev_tracks = [ap.Track.from_host(ctx, s, ch, sample_rate) for s, ch in tracks]
        self.g = ap.Conductor(ctx, out_channels, sample_rate, self.dev_tracks)

    def both(self, name, *a, **k):
        """apply to both; error codes must agree (REF_PANIC on both or success on both)"""
        eo = eg = None
        try:
            getattr(self.o, name)(*self._conv(a, oracle), **k)
        except oracle.OracleError as e:
            eo = e.code
        try:
            getattr(self.g, name)(*self._conv(a, ap), **k)
        except blast.BlastError as e:
            eg = e.code
        assert eo == eg, (name, a, eo, eg)
        return eo

    @staticmethod
    def _conv(args, mod):
        out = []
        for x in args:
            if isinstance(x, dict) and x.get("_tempo"):
                out.append(mod.tempo_repr(x["idx"], x["owned"], x["mode"], x["unit"], x["interval"]))
            else:
                out.append(x)
        return tuple(out)

    def coordinate(self, frames):
        a = self.o.coordinate(frames)
        b = self.g.coordinate(frames)
        assert a.shape == b.shape
        if not np.array_equal(a, b):
            bad = np.flatnonzero(a != b)
            raise AssertionError(f"bus differs at {bad[:8]} (of {bad.size}): oracle {a[bad[:8]]} gpu {b[bad[:8]]}")
        self.same_state()
        return a

    def same_state(self):
        assert self.o.n_groups() == self.g.n_groups()
        for grp in range(-1, self.o.n_groups()):
            n = self.o.n_voices(grp)
            assert n == self.g.n_voices(grp)
            for i in range(n):
                a, b = self.o.get_voice(i, grp), self.g.get_voice(i, grp)
                for f in ("active", "velocity", "gain", "end", "channels", "tempo_current", "tempo_active"):
                    assert getattr(a, f) == getattr(b, f), (grp, i, f, getattr(a, f), getattr(b, f))
                pa, pb = np.float32(a.position), np.float32(b.position)
                assert pa.view(np.uint32) == pb.view(np.uint32) or pa == pb, (grp, i, "position", pa, pb)
        assert self.o.clock() == self.g.clock()


def T(idx=0, owned=True, mode=ap.TM_TBD, unit=ap.TU_SAMPLES, interval=0.0):
    return dict(_tempo=True, idx=idx, owned=owned, mode=mode, unit=unit, interval=interval)


def rng_state(seed):
    r = oracle.Rng(seed)
    return r.state


def test_plain_transport(ctx):
    p = Pair(ctx, 2, make_tracks(1, [(5000, 2), (4000, 1), (6000, 2)]))
    for t in range(3):
        p.both("load", t)
    p.coordinate(100)                       # nothing active: silence
    p.both("start", 0)
    p.coordinate(700)
    p.both("start", 1)
    p.both("velocity", 2, 0.37)
    p.both("start", 2)
    p.coordinate(1500)
    p.both("pause", 0)
    p.coordinate(300)
    p.both("resume", 0)
    p.both("velocity", 1, -1.0)             # takes effect on the running voice: walks backwards
    p.coordinate(900)
    p.both("stop", 2)
    p.both("start", 1)                      # negative velocity: starts at `end`, silent forever (engine.rs:340-343)
    p.coordinate(2100)
    p.g.set_voice(0, gain=0.5)
    p.o.set_voice(0, gain=0.5)
    p.coordinate(4000)                      # voice 0 runs off its end and stalls


@pytest.mark.parametrize("oc", [1, 2, 4])
def test_seq_retrigger_own_tempo(ctx, oc):
    """default Load tempo (owned, TBD, interval = sample rate): overridden with a short Samples tempo so that the
    Seq fires every few hundred calls; chance 100 / 50 / 0"""
    p = Pair(ctx, oc, make_tracks(2, [(30000, 2), (30000, 1), (30000, 2)]))
    for t in range(3):
        p.both("load", t, T(mode=ap.TM_VOICE, interval=64.0))
    p.both("seq", 0, T(owned=False, mode=ap.TM_VOICE, idx=0), 4, [0.0, 1.0, 2.5, 3.0], [100.0, 50.0, 100.0, 0.0], rng_state(7))
    p.both("seq", 1, T(owned=False, mode=ap.TM_VOICE, idx=1), 3, [0.0, 2.0], [80.0, 30.0], rng_state(8))
    p.both("velocity", 2, 1.25)
    p.both("seq", 2, T(owned=False, mode=ap.TM_VOICE, idx=2), 5, [1.0], [60.0], rng_state(9))
    for v in range(3):
        p.both("start", v)
    p.coordinate(3000)
    p.coordinate(1)
    p.coordinate(4097)
    p.both("velocity", 0, 0.5)
    p.coordinate(6000)


def test_seq_process_tempo_and_shared_tempi(ctx):
    """Seq with its own Process tempo (ticked through proc_tempi), a Seq borrowing ANOTHER voice's tempo (sees that
    voice's tick of the same call), a tempo shared by two voices (ticks twice per call) and a context tempo that
    nobody ticks (blast_time.rs:51 vs engine.rs:46-81)."""
    p = Pair(ctx, 2, make_tracks(3, [(20000, 2)] * 4))
    p.both("tc", T(mode=ap.TM_CONTEXT, interval=10.0))
    p.both("load", 0, T(mode=ap.TM_VOICE, interval=50.0))
    p.both("load", 1, T(owned=False, mode=ap.TM_VOICE, idx=0))          # shares voice 0's tempo: both tick it
    p.both("load", 2)                                                    # default TBD tempo, interval = SR
    p.both("load", 3, T(mode=ap.TM_VOICE, unit=ap.TU_MILLIS, interval=2.0))   # 96 samples
    p.both("seq", 0, T(mode=ap.TM_PROCESS, interval=100.0), 2, [0.0, 1.0], [100.0, 100.0], rng_state(1))
    p.both("seq", 1, T(owned=False, mode=ap.TM_VOICE, idx=0), 4, [1.0, 3.0], [100.0, 70.0], rng_state(2))
    p.both("seq", 2, T(owned=False, mode=ap.TM_VOICE, idx=3), 2, [0.0], [100.0], rng_state(3))   # borrows voice 3's
    p.both("seq", 3, T(owned=False, mode=ap.TM_CONTEXT, idx=0), 1, [0.0], [100.0], rng_state(4))
    p.both("seq", 3, T(owned=False, mode=ap.TM_VOICE, idx=3), 8, [0.0, 4.0], [50.0, 50.0], rng_state(5))
    for v in range(4):
        p.both("start", v)
    p.coordinate(2500)
    p.both("start", 0, ap.IDX_TEMPO)        # context tempo now active and stuck at 0: the Seq on voice 3 fires on
    p.coordinate(40)                        # every call until ... forever (steps = [0.0], current stays 0)
    p.both("stop", 0, ap.IDX_TEMPO)
    p.coordinate(3000)
    p.both("pause", 0)                      # voice 0 stops ticking the shared tempo: rate drops to 1
    p.coordinate(2000)
    p.both("stop", 1)
    p.both("start", 1)
    p.coordinate(2000)


def test_groups(ctx):
    p = Pair(ctx, 2, make_tracks(4, [(15000, 2), (15000, 1), (15000, 2), (15000, 2)]))
    for t in range(4):
        p.both("load", t, T(mode=ap.TM_VOICE, interval=32.0))
    p.both("seq", 0, T(owned=False, mode=ap.TM_VOICE, idx=0), 4, [0.0, 2.0], [100.0, 100.0], rng_state(11))
    p.both("seq", 2, T(mode=ap.TM_PROCESS, interval=77.0), 3, [1.0], [90.0], rng_state(12))
    # group of voices 2 (keeps its tempo) and 0 (adopts the group tempo, and so does its process 0)
    p.both("group", T(mode=ap.TM_GROUP, unit=ap.TU_BPM, interval=48000.0), [(2, False, []), (0, True, [0])])
    assert p.g.n_voices() == 2 and p.g.n_voices(0) == 2
    p.both("seq", 0, T(owned=False, mode=ap.TM_GROUP, idx=0), 2, [0.0], [100.0], rng_state(13), ap.IDX_GROUP)   # stored, never run
    p.both("start", 0)                      # Conductor.voices[0] is now the old voice 1
    p.coordinate(1000)
    p.both("start", 0, ap.IDX_GROUP)
    p.coordinate(3000)
    p.both("pause", 0, ap.IDX_GROUP)
    p.coordinate(500)
    p.both("resume", 0, ap.IDX_GROUP)
    p.coordinate(2500)
    p.both("stop", 0, ap.IDX_GROUP)         # voices inactive but NOT rewound (engine.rs:516-528)
    p.coordinate(200)
    p.both("start", 0, ap.IDX_GROUP)
    p.both("unload", 0)
    p.coordinate(1800)


def test_seq_bpm_interval_rare_hits(ctx):
    """a BPM tempo gives a non-integer interval: hits need current/interval to ROUND to the exact step value"""
    p = Pair(ctx, 2, make_tracks(5, [(60000, 2), (60000, 2)]))
    p.both("load", 0, T(mode=ap.TM_VOICE, unit=ap.TU_BPM, interval=7000.0))      # 411.43 samples
    p.both("load", 1, T(mode=ap.TM_VOICE, unit=ap.TU_BPM, interval=9000.0))      # exactly 320
    for v in range(2):
        p.both("seq", v, T(owned=False, mode=ap.TM_VOICE, idx=v), 4, [0.0, 1.0, 2.0, 3.0], [100.0] * 4, rng_state(20 + v))
        p.both("start", v)
    p.coordinate(50000)


def test_many_retriggers_forces_chunking(ctx):
    """a retrigger every 8 calls: far more than kMaxEvents / kMaxSeg per span -> the span is cut into chunks"""
    p = Pair(ctx, 2, make_tracks(6, [(9000, 2), (9000, 1)]))
    p.both("load", 0, T(mode=ap.TM_VOICE, interval=8.0))
    p.both("load", 1, T(mode=ap.TM_VOICE, interval=1.0))
    p.both("seq", 0, T(owned=False, mode=ap.TM_VOICE, idx=0), 1, [0.0], [75.0], rng_state(31))
    p.both("seq", 1, T(owned=False, mode=ap.TM_VOICE, idx=1), 16, [3.0, 9.0], [100.0, 40.0], rng_state(32))
    p.both("velocity", 0, 0.8)
    p.both("start", 0)
    p.both("start", 1)
    p.coordinate(20000)


def test_ref_panics_leave_state_untouched(ctx):
    p = Pair(ctx, 2, make_tracks(8, [(3000, 2), (3000, 2)]))
    assert p.both("load", 5) == _lib.ERR_REF_PANIC
    p.both("load", 0)
    p.both("load", 1)
    assert p.both("start", 2) == _lib.ERR_REF_PANIC
    assert p.both("velocity", 9, 2.0) == _lib.ERR_REF_PANIC
    assert p.both("unload", 2) == _lib.ERR_REF_PANIC
    assert p.both("load", 0, T(owned=False, mode=ap.TM_GROUP, idx=0)) == _lib.ERR_REF_PANIC
    assert p.both("start", 0, ap.IDX_GROUP) == _lib.ERR_REF_PANIC
    assert p.both("seq", 4, T(), 4, [0.0], [100.0], rng_state(1)) == _lib.ERR_REF_PANIC
    # GPU side validates the whole member list before moving anything
    with pytest.raises(blast.BlastError) as e:
        p.g.group(ap.tempo_repr(), [(0, False, []), (1, False, [])])     # after removing 0 only index 0 is left
    assert e.value.code == _lib.ERR_REF_PANIC
    assert p.g.n_voices() == 2 and p.g.n_groups() == 0
    p.both("start", 0)
    p.both("start", 1)
    p.coordinate(500)
    # a Seq with an empty step list panics when it is first processed with an active tempo (processes.rs:79)
    p.both("seq", 0, T(owned=False, mode=ap.TM_VOICE, idx=0), 4, [], [], rng_state(2))
    with pytest.raises(oracle.OracleError) as eo:
        p.o.coordinate(10)
    with pytest.raises(blast.BlastError) as eg:
        p.g.coordinate(10)
    assert eo.value.code == oracle.REF_PANIC and eg.value.code == _lib.ERR_REF_PANIC


def test_timeline_matches_stepwise_oracle(ctx):
    tracks = make_tracks(9, [(40000, 2), (40000, 1), (40000, 2)])
    o = oracle.Conductor(2, SR, [(s, ch, SR) for s, ch in tracks])
    dev = [ap.Track.from_host(ctx, s, ch, SR) for s, ch in tracks]
    g = ap.Conductor(ctx, 2, SR, dev)

    def tr(mod, **k):
        return mod.tempo_repr(**k)

    script = [
        (0, "load", lambda m: (0, tr(m, mode=ap.TM_VOICE, interval=128.0))),
        (0, "load", lambda m: (1,)),
        (0, "load", lambda m: (2, tr(m, mode=ap.TM_VOICE, interval=100.0))),
        (0, "seq", lambda m: (0, tr(m, owned=False, mode=ap.TM_VOICE, idx=0), 2, [0.0, 1.0], [100.0, 35.0], rng_state(51))),
        (0, "start", lambda m: (0,)),
        (1000, "start", lambda m: (1,)),
        (1000, "velocity", lambda m: (1, 0.61)),
        (4097, "start", lambda m: (2,)),
        (9000, "seq", lambda m: (2, tr(m, mode=ap.TM_PROCESS, interval=333.0), 3, [2.0], [100.0], rng_state(52))),
        (9000, "stop", lambda m: (2,)),
        (9001, "start", lambda m: (2,)),
        (15000, "pause", lambda m: (0,)),
        (15000, "unload", lambda m: (1,)),
        (20000, "resume", lambda m: (0,)),
        (26000, "velocity", lambda m: (0, 2.5)),
    ]
    total = 30000
    # oracle: step by step
    out = []
    cur = 0
    for frame, name, args in script:
        if frame > cur:
            out.append(o.coordinate(frame - cur))
            cur = frame
        getattr(o, name)(*args(oracle))
    out.append(o.coordinate(total - cur))
    expect = np.concatenate(out)
    # GPU: one timeline call
    builders = {"load": ap.Cmd.load, "seq": ap.Cmd.seq, "velocity": ap.Cmd.velocity, "unload": ap.Cmd.unload,
                "start": lambda i: ap.Cmd.transport(ap.CMD_START, i), "stop": lambda i: ap.Cmd.transport(ap.CMD_STOP, i),
                "pause": lambda i: ap.Cmd.transport(ap.CMD_PAUSE, i), "resume": lambda i: ap.Cmd.transport(ap.CMD_RESUME, i)}
    timeline = [(frame, builders[name](*args(ap))) for frame, name, args in script]
    got = g.render_timeline(timeline, total)
    assert np.array_equal(got, expect)
    assert g.clock() == o.clock() == total
    with pytest.raises(blast.BlastError):
        g.render_timeline([(10, ap.Cmd.quit()), (5, ap.Cmd.quit())], 20)      # unsorted


def test_sharded_partials_sum_to_the_whole(ctx):
    tracks = make_tracks(10, [(12000, 2)] * 5)
    dev = [ap.Track.from_host(ctx, s, ch, SR) for s, ch in tracks]
    frames = 9000

    def build(rank, world):
        g = ap.Conductor(ctx, 2, SR, dev)
        g.set_shard(rank, world)
        for t in range(5):
            g.load(t, ap.tempo_repr(mode=ap.TM_VOICE, interval=40.0 + t))
            g.seq(t, ap.tempo_repr(owned=False, mode=ap.TM_VOICE, idx=0), 2, [0.0], [100.0], rng_state(60 + t))
            g.velocity(t, 0.5 + 0.25 * t)
            g.start(t)
        return g

    n = frames * 2
    parts = []
    for rank in range(2):
        g = build(rank, 2)
        buf = ctx.alloc(4 * n)
        g.render_partial_dev(frames, buf.ptr)
        parts.append(buf.download(np.int32, n))
    whole = build(0, 1).coordinate(frames)
    summed = (parts[0].astype(np.int64) + parts[1].astype(np.int64)).astype(np.int16)   # low 16 bits
    assert np.array_equal(summed, whole)
    o = oracle.Conductor(2, SR, [(s, ch, SR) for s, ch in tracks])
    for t in range(5):
        o.load(t, oracle.tempo_repr(mode=oracle.TM_VOICE, interval=40.0 + t))
        o.seq(t, oracle.tempo_repr(owned=False, mode=oracle.TM_VOICE, idx=0), 2, [0.0], [100.0], rng_state(60 + t))
        o.velocity(t, 0.5 + 0.25 * t)
        o.start(t)
    assert np.array_equal(o.coordinate(frames), whole)


import os as _os


@pytest.mark.parametrize("seed", range(int(_os.environ.get("BLAST_FUZZ_SEEDS", "6"))))
def test_random_command_streams(ctx, seed):
    """fuzz: random commands (valid and invalid indices), random spans, everything compared after every span"""
    r = np.random.default_rng(1000 + seed)
    oc = int(r.choice([1, 2, 2, 3]))
    tracks = make_tracks(100 + seed, [(int(r.integers(200, 9000)), int(r.choice([1, 2, 2, 3]))) for _ in range(4)])
    p = Pair(ctx, oc, tracks)
    intervals = [8.0, 25.0, 64.0, 100.0, 333.0, 48000.0 / 7]
    for step in range(40):
        nv, ng = p.o.n_voices(), p.o.n_groups()
        kind = r.choice(["load", "start", "start", "pause", "resume", "stop", "velocity", "seq", "seq", "group", "tc",
                         "unload", "gstart", "gstop", "render", "render", "render"])
        def rt():
            m = int(r.choice([ap.TM_VOICE, ap.TM_TBD, ap.TM_PROCESS, ap.TM_GROUP, ap.TM_CONTEXT]))
            owned = bool(r.random() < 0.6)
            return T(idx=int(r.integers(0, 3)), owned=owned, mode=m, unit=int(r.integers(0, 3)) if owned and r.random() < 0.3 else 0,
                     interval=float(r.choice(intervals)))
        if kind == "load":
            p.both("load", int(r.integers(0, 5)), rt())
        elif kind in ("start", "pause", "resume", "stop"):
            p.both(kind, int(r.integers(0, nv + 1)))
        elif kind == "velocity":
            p.both("velocity", int(r.integers(0, nv + 1)), float(r.choice([1.0, 0.5, 1.5, -1.0, 0.123, 2.0, 0.0])))
        elif kind == "seq":
            n = int(r.integers(1, 4))
            steps = [float(x) for x in r.integers(0, 4, size=n)]
            chance = [float(x) for x in r.choice([0.0, 30.0, 100.0], size=n)]
            p.both("seq", int(r.integers(0, nv + 1)), rt(), int(r.integers(1, 6)), steps, chance, rng_state(int(r.integers(0, 1 << 30))))
        elif kind == "group" and nv > 0:
            k = int(r.integers(1, min(nv, 2) + 1))
            members = []
            left = nv
            for _ in range(k):
                members.append((int(r.integers(0, left)), bool(r.random() < 0.5), []))
                left -= 1
            p.both("group", rt(), members)
        elif kind == "tc":
            p.both("tc", T(mode=ap.TM_CONTEXT, interval=float(r.choice(intervals))))
        elif kind == "unload":
            p.both("unload", int(r.integers(0, nv + 1)))
        elif kind == "gstart":
            p.both("start", int(r.integers(0, ng + 1)), ap.IDX_GROUP)
        elif kind == "gstop":
            p.both(str(r.choice(["stop", "pause", "resume"])), int(r.integers(0, ng + 1)), ap.IDX_GROUP)
        else:
            p.coordinate(int(r.choice([1, 7, 300, 2048, 2500, 5000])))
    p.coordinate(3000)


def test_long_span_many_epochs_overflow_path(ctx):
    """one coordinate() call that needs hundreds of position segments per voice: the first attempts overflow the segment
    list on the device (K4 must then skip the void render instead of walking truncated trajectories), the span is
    re-run in halves until it fits"""
    p = Pair(ctx, 2, make_tracks(12, [(400000, 2)] * 3 + [(400000, 1)]))
    for t in range(4):
        p.both("load", t, T(mode=ap.TM_VOICE, interval=float(1500 + 100 * t)))
        p.both("seq", t, T(owned=False, mode=ap.TM_VOICE, idx=t), 4, [0.0, 1.0, 2.0, 3.0], [50.0, 100.0, 25.0, 75.0], rng_state(90 + t))
        p.both("velocity", t, [1.0, 0.77, 1.31, 1.0][t])
        p.both("start", t)
    p.coordinate(300000)
    p.coordinate(50000)


@pytest.mark.parametrize("seed", range(int(_os.environ.get("BLAST_FUZZ_SEEDS", "6"))))
def test_random_degenerate_tempi_and_seqs(ctx, seed):
    """fuzz over the corners of TempoState / Seq arithmetic: zero, negative, tiny, infinite and NaN intervals (via every
    unit), period 0, steps that are negative / fractional / -0.0 / beyond the period, chances outside [0, 100] and NaN
    (`chance as i64` saturates, NaN -> 0), tempi ticked by nobody or by several voices"""
    r = np.random.default_rng(5000 + seed)
    oc = int(r.choice([1, 2, 2, 4]))
    tracks = make_tracks(300 + seed, [(int(r.integers(50, 3000)), int(r.choice([1, 2, 2, 3]))) for _ in range(3)])
    p = Pair(ctx, oc, tracks)
    weird_iv = [0.0, -5.0, 1e-3, 0.5, 1.0, 2.0, 7.5, float("inf"), float("nan"), 1e30, 3.0]
    weird_steps = [0.0, -0.0, 1.0, 0.5, -1.0, 2.0, 3.0, 7.0, float("nan"), 0.25]
    weird_chance = [0.0, 100.0, 50.0, -5.0, 100.5, float("nan"), 1e30, 99.99]
    p.both("tc", T(mode=ap.TM_CONTEXT, unit=int(r.integers(0, 3)), interval=float(r.choice(weird_iv))))
    for t in range(3):
        owned = bool(r.random() < 0.7) or t == 0
        p.both("load", t, T(idx=0, owned=owned, mode=ap.TM_VOICE if owned else int(r.choice([ap.TM_VOICE, ap.TM_CONTEXT])),
                            unit=int(r.integers(0, 3)), interval=float(r.choice(weird_iv))))
    for _ in range(int(r.integers(2, 7))):
        n = int(r.integers(1, 4))
        mode = int(r.choice([ap.TM_VOICE, ap.TM_PROCESS, ap.TM_CONTEXT]))
        owned = mode == ap.TM_PROCESS or bool(r.random() < 0.3)
        p.both("seq", int(r.integers(0, 3)), T(idx=int(r.integers(0, 3)) if mode == ap.TM_VOICE else 0, owned=owned, mode=mode,
                                               unit=int(r.integers(0, 3)), interval=float(r.choice(weird_iv))),
               int(r.choice([0, 1, 2, 4, 4, 16])), [float(x) for x in r.choice(weird_steps, size=n)],
               [float(x) for x in r.choice(weird_chance, size=n)], rng_state(int(r.integers(0, 1 << 30))))
    for v in range(3):
        p.both("velocity", v, float(r.choice([1.0, 0.5, 1.25, -1.0, 0.0, 3.0])))
        p.both("start", v)
    if r.random() < 0.5:
        p.both("start", 0, ap.IDX_TEMPO)
    for _ in range(4):
        p.coordinate(int(r.choice([1, 3, 64, 300, 1500])))
        if r.random() < 0.3:
            p.both(str(r.choice(["stop", "start", "pause", "resume"])), int(r.integers(0, 3)))


def test_retrigger_on_every_call(ctx):
    """a context tempo nobody ticks stays at 0, so a Seq with step -0.0 fires on EVERY call: the voice is reset before
    each channel's read and still advances once per frame (position 1.0 after every frame).  Regression: the epoch
    template must also cover the position after an epoch's last advance."""
    p = Pair(ctx, 2, make_tracks(13, [(500, 2), (400, 1)]))
    p.both("tc", T(mode=ap.TM_CONTEXT, interval=2.0))
    p.both("start", 0, ap.IDX_TEMPO)
    for t in range(2):
        p.both("load", t)
        p.both("seq", t, T(owned=False, mode=ap.TM_CONTEXT, idx=0), 4, [-0.0], [100.0], rng_state(70 + t))
        p.both("start", t)
    for frames in (1, 1, 3, 64, 300):
        p.coordinate(frames)
    assert p.g.get_voice(0).position == 1.0 and p.g.get_voice(1).position == 1.0
