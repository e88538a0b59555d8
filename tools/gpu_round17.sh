#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conductor_gpu.py -x -q > gpurun_out/pytest17.log 2>&1; echo "pytest_rc=$?"; tail -15 gpurun_out/pytest17.log
timeout 400 python tools/bench_render.py --skip-c4 > gpurun_out/bench_render17.json 2> gpurun_out/bench_render17.err; echo "render_rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_render17.json')); print({k:(v.get('ms'),v.get('GBps')) for k,v in d.items() if 'ms' in v})"; tail -3 gpurun_out/bench_render17.err
