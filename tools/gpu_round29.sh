#!/bin/bash
mkdir -p gpurun_out
BLAST_FUZZ_TRIALS=600 timeout 1200 python -m pytest tests/test_mpeg_gpu.py -x -q -k random_buffers > gpurun_out/pytest29.log 2>&1; echo "pytest_rc=$?"; tail -8 gpurun_out/pytest29.log
