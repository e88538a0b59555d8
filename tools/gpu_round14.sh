#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_rng_gpu.py -x -q > gpurun_out/pytest14.log 2>&1; echo "pytest_rc=$?"; tail -3 gpurun_out/pytest14.log
timeout 200 python tools/bench_render.py > gpurun_out/bench_render14.json 2> gpurun_out/bench_render14.err; echo "render_rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_render14.json')); print(d.get('c4'))"
