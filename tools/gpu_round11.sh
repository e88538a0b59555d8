#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest11.log 2>&1; echo "pytest_rc=$?" | tee -a gpurun_out/pytest11.log
tail -4 gpurun_out/pytest11.log
timeout 600 python bench.py > gpurun_out/bench_r11.json 2> gpurun_out/bench_r11.err; echo "bench_rc=$?"; tail -3 gpurun_out/bench_r11.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r11.json 2> gpurun_out/bench_ref_r11.err; echo "ref_rc=$?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke11.log 2>&1; echo "smoke_rc=$?"; tail -2 gpurun_out/smoke11.log
