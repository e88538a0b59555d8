#!/usr/bin/env python
"""C1 on the CPU: the faithful oracle restatement of wav::parse (per-pair bounds-checked reads, Vec growth) on one
synthetic 10-minute 16-bit stereo 44.1 kHz WAV, single thread like the reference.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402
import synth  # noqa: E402

if __name__ == "__main__":
    img = synth.wav_image(0xC1, synth.C1_DATA_LEN)
    oracle.wav_parse(synth.wav_image(1, 4096))                    # warm the library
    t0 = time.perf_counter()
    _, got = oracle.wav_parse(img)
    dt = time.perf_counter() - t0
    assert np.array_equal(got, img[44:].view("<i2"))
    print(json.dumps({"cpu_oracle_faithful_1_thread": {"ms": round(dt * 1e3, 1), "gsamples_per_s": round(got.size / dt / 1e9, 3)}}))
