#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_mpeg_gpu.py -x -q > gpurun_out/pytest4.log 2>&1; echo "pytest_rc=$?" | tee -a gpurun_out/pytest4.log
tail -5 gpurun_out/pytest4.log
timeout 300 python tools/bench_mpeg.py > gpurun_out/bench_mpeg.json 2> gpurun_out/bench_mpeg.err; echo "bench_rc=$?"
cat gpurun_out/bench_mpeg.json; tail -5 gpurun_out/bench_mpeg.err
CMD="python tools/bench_mpeg.py --gib 4 --iters 1"
timeout 300 $CMD > gpurun_out/plain4.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:mpeg_sync_scan -s 1 -c 1 -o gpurun_out/prof_mpeg $CMD > gpurun_out/ncu_mpeg.log 2>&1
tail -2 gpurun_out/ncu_mpeg.log
