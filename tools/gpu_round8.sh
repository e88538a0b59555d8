#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/bench_mpeg.py --gib 4 --iters 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mpeg_sync_scan -s 1 -c 1 -f -o gpurun_out/prof_mpeg_v3 $CMD > gpurun_out/ncu_mpeg_v3.log 2>&1
tail -2 gpurun_out/ncu_mpeg_v3.log
