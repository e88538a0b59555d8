#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_render_gpu.py tests/test_conductor_gpu.py tests/test_rng_gpu.py -x -q > gpurun_out/pytest13.log 2>&1; echo "pytest_rc=$?" | tee -a gpurun_out/pytest13.log
tail -4 gpurun_out/pytest13.log
timeout 200 python tools/bench_render.py > gpurun_out/bench_render13.json 2> gpurun_out/bench_render13.err; echo "render_rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_render13.json'))
print({k:(v.get('ms'),v.get('GBps')) for k,v in d.items() if 'ms' in v}); print(d.get('c4'))"
timeout 300 python bench.py --no-e2e --no-cpu > gpurun_out/bench_r13.json 2>gpurun_out/bench_r13.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r13.json')); print(d['value'], d['kernel_ms'], d['gpu_launches'])"
