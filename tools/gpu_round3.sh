#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_render_gpu.py -x -q > gpurun_out/pytest3.log 2>&1; echo "pytest_rc=$?" | tee -a gpurun_out/pytest3.log
tail -15 gpurun_out/pytest3.log
timeout 600 python tools/bench_render.py --skip-c4 > gpurun_out/bench_render_tma.json 2> gpurun_out/bench_render_tma.err; echo "bench_rc=$?"
cat gpurun_out/bench_render_tma.json; tail -5 gpurun_out/bench_render_tma.err
CMD="python tools/bench_render.py --skip-c4 --iters 1"
$CMD > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:voice_render_mix_tma -s 2 -c 2 -o gpurun_out/prof_render_tma2 $CMD > gpurun_out/ncu_render.log 2>&1
tail -3 gpurun_out/ncu_render.log
