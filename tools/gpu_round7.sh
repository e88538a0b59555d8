#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_mpeg_gpu.py -x -q > gpurun_out/pytest7.log 2>&1; echo "pytest_rc=$?" | tee -a gpurun_out/pytest7.log
tail -15 gpurun_out/pytest7.log
timeout 200 python tools/bench_mpeg.py > gpurun_out/bench_mpeg_v3.json 2> gpurun_out/bench_mpeg_v3.err; echo "bench_mpeg_rc=$?"
cat gpurun_out/bench_mpeg_v3.json; tail -3 gpurun_out/bench_mpeg_v3.err
