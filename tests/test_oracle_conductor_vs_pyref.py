"""The C++ oracle's Conductor (commands, shared tempi, groups, Seq) against the independent pure-Python restatement in
tests/pyref.py, on random command streams.  CPU only; small sizes (pure-Python loops)."""
import numpy as np
import pytest

import oracle
import pyref

KIND = {"voice": oracle.IDX_VOICE, "group": oracle.IDX_GROUP, "tempo": oracle.IDX_TEMPO}


class Both:
    def __init__(self, oc, sr, tracks):
        self.o = oracle.Conductor(oc, sr, [(s, ch, sr) for s, ch in tracks])
        self.p = pyref.PyConductor(oc, sr, tracks)

    def do(self, name, *a, kind=None):
        eo = ep = False
        oa = list(a)
        for i, x in enumerate(oa):
            if isinstance(x, tuple) and len(x) == 5 and isinstance(x[1], bool):
                oa[i] = oracle.tempo_repr(*x)
        try:
            if kind is None:
                getattr(self.o, name)(*oa)
            else:
                getattr(self.o, name)(*oa, idx_kind=KIND[kind])
        except oracle.OracleError as e:
            assert e.code == oracle.REF_PANIC
            eo = True
        try:
            if kind is None:
                getattr(self.p, name)(*a)
            else:
                getattr(self.p, name)(*a, idx_kind=kind)
        except pyref.RefPanic:
            ep = True
        assert eo == ep, (name, a, eo, ep)
        return eo

    def coordinate(self, frames):
        a, b = self.o.coordinate(frames), self.p.coordinate(frames)
        assert np.array_equal(a, b), np.flatnonzero(a != b)[:8]
        voices = list(self.p.voices) + [v for g in self.p.groups for v in g.voices]
        ov = [self.o.get_voice(i) for i in range(self.o.n_voices())] + \
             [self.o.get_voice(i, g) for g in range(self.o.n_groups()) for i in range(self.o.n_voices(g))]
        assert len(voices) == len(ov)
        for pv, v in zip(voices, ov):
            assert bool(v.active) == pv.active and v.tempo_current == pv.tempo.current and bool(v.tempo_active) == pv.tempo.active
            pa, pb = np.float32(v.position), np.float32(pv.position)
            assert pa == pb or (np.isnan(pa) and np.isnan(pb)), (pa, pb)


@pytest.mark.parametrize("seed", range(12))
def test_random_command_streams(seed):
    r = np.random.default_rng(7000 + seed)
    oc = int(r.choice([1, 2, 2, 3]))
    tracks = [(r.integers(-20000, 20000, size=int(r.integers(20, 400)) * ch).astype(np.int16), ch)
              for ch in (int(r.choice([1, 2, 2, 3])) for _ in range(3))]
    b = Both(oc, 48000, tracks)
    ivs = [1.0, 2.0, 3.0, 7.5, 0.5, 16.0, 0.0, float("inf")]

    def rt():
        mode = int(r.choice([pyref.TM_VOICE, pyref.TM_TBD, pyref.TM_PROCESS, pyref.TM_GROUP, pyref.TM_CONTEXT]))
        owned = bool(r.random() < 0.6)
        return (int(r.integers(0, 3)), owned, mode, int(r.integers(0, 3)) if owned and r.random() < 0.3 else 0, float(r.choice(ivs)))

    for _ in range(30):
        nv, ng = len(b.p.voices), len(b.p.groups)
        k = r.choice(["load", "start", "start", "stop", "pause", "resume", "velocity", "seq", "seq", "group", "tc", "unload",
                      "gstart", "gstop", "tstart", "render", "render"])
        if k == "load":
            b.do("load", int(r.integers(0, 4)), rt())
        elif k in ("start", "stop", "pause", "resume"):
            b.do(k, int(r.integers(0, nv + 1)), kind="voice")
        elif k == "velocity":
            b.do("velocity", int(r.integers(0, nv + 1)), float(r.choice([1.0, 0.5, 1.5, -1.0, 0.0])))
        elif k == "seq":
            n = int(r.integers(1, 4))
            st = oracle.Rng(int(r.integers(0, 1 << 30))).state
            b.do("seq", int(r.integers(0, nv + 1)), rt(), int(r.integers(0, 5)), [float(x) for x in r.integers(0, 4, size=n)],
                 [float(x) for x in r.choice([0.0, 40.0, 100.0], size=n)], st, kind="voice")
        elif k == "group" and nv > 0:
            members, left = [], nv
            for _m in range(int(r.integers(1, min(nv, 2) + 1))):
                members.append((int(r.integers(0, left)), bool(r.random() < 0.5), []))
                left -= 1
            b.do("group", rt(), members)
        elif k == "tc":
            b.do("tc", (0, True, pyref.TM_CONTEXT, 0, float(r.choice(ivs))))
        elif k == "unload":
            b.do("unload", int(r.integers(0, nv + 1)))
        elif k == "gstart":
            b.do("start", int(r.integers(0, ng + 1)), kind="group")
        elif k == "gstop":
            b.do(str(r.choice(["stop", "pause", "resume"])), int(r.integers(0, ng + 1)), kind="group")
        elif k == "tstart":
            b.do(str(r.choice(["start", "stop", "pause", "resume"])), int(r.integers(0, len(b.p.tempo_cons) + 1)), kind="tempo")
        else:
            b.coordinate(int(r.choice([1, 5, 40, 120])))
    b.coordinate(60)
