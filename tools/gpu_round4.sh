#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_mpeg_gpu.py -x -q > gpurun_out/pytest4.log 2>&1; echo "pytest_rc=$?" | tee -a gpurun_out/pytest4.log
tail -25 gpurun_out/pytest4.log
timeout 300 python tools/bench_mpeg.py > gpurun_out/bench_mpeg.json 2> gpurun_out/bench_mpeg.err; echo "bench_rc=$?"
cat gpurun_out/bench_mpeg.json; tail -5 gpurun_out/bench_mpeg.err
