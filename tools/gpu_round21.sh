#!/bin/bash
mkdir -p gpurun_out
BLAST_CONDUCTOR_DEBUG=1 timeout 300 python tools/bench_render.py --skip-c4 --only-seq 2>&1 | tail -9
bash tools/gpu_round20.sh
