#!/bin/bash
mkdir -p gpurun_out
BLAST_FUZZ_SEEDS=400 timeout 1200 python -m pytest tests/test_render_gpu.py -x -q -k weird > gpurun_out/pytest28.log 2>&1; echo "pytest_rc=$?"; tail -25 gpurun_out/pytest28.log
