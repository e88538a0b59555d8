#!/bin/bash
mkdir -p gpurun_out
BLAST_FUZZ_SEEDS=60 timeout 600 python -m pytest tests/test_render_gpu.py tests/test_conductor_gpu.py tests/test_golden_gpu.py -x -q > gpurun_out/pytest19.log 2>&1; echo "pytest_rc=$?"; tail -4 gpurun_out/pytest19.log
BLAST_CONDUCTOR_DEBUG=1 timeout 300 python tools/bench_render.py --skip-c4 --only-seq 2>&1 | grep -v "^\[blast" | tail -9
BLAST_CONDUCTOR_DEBUG=1 timeout 300 python tools/bench_render.py --skip-c4 --only-seq 2>&1 | grep "^\[blast" | tail -1
