#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_render_gpu.py -x -q --durations=5 > gpurun_out/pytest26.log 2>&1; echo "pytest_rc=$?"; tail -14 gpurun_out/pytest26.log
free -g | head -2
