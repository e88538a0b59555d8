#!/bin/bash
mkdir -p gpurun_out
BLAST_FUZZ_SEEDS=40 timeout 600 python -m pytest tests/test_render_gpu.py tests/test_conductor_gpu.py tests/test_golden_gpu.py -x -q > gpurun_out/pytest22.log 2>&1; echo "pytest_rc=$?"; tail -4 gpurun_out/pytest22.log
timeout 300 python tools/bench_render.py --skip-c4 2>&1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:(v.get('ms'),v.get('GBps')) for k,v in d.items() if 'ms' in v})"
timeout 300 python bench.py --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['kernel_ms'])"
