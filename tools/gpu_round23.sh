#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decode_gpu.py tests/test_golden_gpu.py -x -q > gpurun_out/pytest23.log 2>&1; echo "pytest_rc=$?"; tail -3 gpurun_out/pytest23.log
timeout 600 python bench.py > gpurun_out/bench_r23.json 2> gpurun_out/bench_r23.err; echo "bench_rc=$?"; tail -3 gpurun_out/bench_r23.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r23.json'))
for k in ('value','kernel_ms','e2e','e2e_parse_dropin'): print(k, {kk:vv for kk,vv in d[k].items() if kk not in ('api','pcie_peak_source')} if isinstance(d[k],dict) else d[k])
PY
