#!/bin/bash
mkdir -p gpurun_out
BLAST_FUZZ_SEEDS=300 timeout 1200 python -m pytest tests/test_conductor_gpu.py -x -q > gpurun_out/pytest27.log 2>&1; echo "pytest_rc=$?"; tail -30 gpurun_out/pytest27.log
