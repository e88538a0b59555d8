import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import audio_decoder_b200 as blast
from test_render_gpu import _random_voice, gpu_render, oracle_render

ctx = blast.Context(0)
for oc in (2, 1):
    rng = np.random.default_rng(7 + oc)
    for trial in range(6):
        voices = [_random_voice(rng, long_clip=True) for _ in range(int(rng.integers(2, 24)))]
        frames = int(rng.integers(4097, 9000))
        bus, pos = gpu_render(ctx, voices, oc, frames)
        exp, epos = oracle_render(voices, oc, frames)
        if np.array_equal(bus, exp):
            continue
        print("MISMATCH oc", oc, "trial", trial, "frames", frames, "first", int(np.argmax(bus != exp)), "count", int((bus != exp).sum()))
        for i, v in enumerate(voices):
            b, _ = gpu_render(ctx, [v], oc, frames)
            e, _ = oracle_render([v], oc, frames)
            if not np.array_equal(b, e):
                bad = np.nonzero(b != e)[0]
                print("  voice", i, "ch", v["channels"], "vel", v["velocity"], "gain", v["gain"], "pos", v["position"], "active", v["active"],
                      "nfr", len(v["samples"]) // v["channels"], "bad", len(bad), "first", bad[:8], "last", bad[-3:], "got", b[bad[:4]], "exp", e[bad[:4]])
