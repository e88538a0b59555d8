#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mpeg_gpu.py tests/test_golden_gpu.py -x -q > gpurun_out/pytest15.log 2>&1; echo "pytest_rc=$?"; tail -12 gpurun_out/pytest15.log
timeout 200 python tools/bench_mpeg.py > gpurun_out/bench_mpeg_v3.json 2> gpurun_out/bench_mpeg_v3.err; echo "bench_mpeg_rc=$?"; head -12 gpurun_out/bench_mpeg_v3.json
