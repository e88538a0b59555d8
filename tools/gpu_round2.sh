#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest2.log 2>&1; echo "pytest_rc=$?" | tee -a gpurun_out/pytest2.log
tail -30 gpurun_out/pytest2.log
timeout 600 python tools/bench_render.py > gpurun_out/bench_render.json 2> gpurun_out/bench_render.err; echo "bench_render_rc=$?"
cat gpurun_out/bench_render.json; tail -5 gpurun_out/bench_render.err
