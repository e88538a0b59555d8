#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/bench_mpeg.py --gib 4 --iters 1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mpeg -c 40 --csv --log-file gpurun_out/mpeg_launches.csv $CMD > /dev/null 2>&1
cut -d, -f5,12- gpurun_out/mpeg_launches.csv | tail -20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mpeg_walk -s 1 -c 1 -f -o gpurun_out/prof_mpeg_walk $CMD > gpurun_out/ncu_mpeg_walk.log 2>&1
tail -2 gpurun_out/ncu_mpeg_walk.log
