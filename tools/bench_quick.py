#!/usr/bin/env python
"""Development timing of selected bench_configs entries on one GPU:  python tools/bench_quick.py c3_unit c3_mixed c2_true24 ..."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audio_decoder_b200 as blast  # noqa: E402
import bench  # noqa: E402
import bench_configs as bc  # noqa: E402

if __name__ == "__main__":
    names = sys.argv[1:] or ["c3_unit", "c3_mixed"]
    peak, _ = bench.measured_peaks()
    out = {}
    with blast.Context(0) as ctx:
        for n in names:
            if n == "c3_unit":
                out[n] = bc.c3(ctx, 0, 1, 0, None, peak, False, True)
            elif n == "c3_mixed":
                out[n] = bc.c3(ctx, 0, 1, 0, None, peak, False, False)
            elif n == "c3_seq":
                out[n] = bc.c3_seq(ctx, peak, False)
            elif n == "c1":
                out[n] = bc.c1(ctx, peak, False)
            elif n == "c2_true24":
                out[n] = bc.c2_true24(ctx, peak, False)
            elif n == "c4":
                out[n] = bc.c4(ctx, peak, False)
            elif n == "c5":
                out[n] = bc.c5(ctx, peak, False)
            ctx.trim()
    for k, v in out.items():
        print(k, json.dumps(v))
