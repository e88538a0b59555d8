#!/bin/bash
# One-GPU validation as the driver runs it at round end: smoke, the gpu-marked parity tests, both bench arms.
#   gpurun --timeout 1800 -- 'bash tools/gpu_validate.sh'
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke_rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench_rc=$?"; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref_rc=$?"
