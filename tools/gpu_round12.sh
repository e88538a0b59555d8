#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
timeout 300 $CMD > gpurun_out/plain12.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches.csv $CMD > gpurun_out/ncu12.log 2>&1
grep -c "gpu__time_duration" gpurun_out/bench_launches.csv
timeout 200 python tools/bench_render.py > gpurun_out/bench_render12.json 2> gpurun_out/bench_render12.err; echo "render_rc=$?"; cat gpurun_out/bench_render12.json | head -60
