#!/bin/bash
mkdir -p gpurun_out
BLAST_MPEG_DEBUG_NOCHAIN=1 timeout 200 python tools/bench_mpeg.py --iters 3 2>&1 | head -12
