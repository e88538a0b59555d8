#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke24.log 2>&1; echo "smoke_rc=$?"; tail -2 gpurun_out/smoke24.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest24.log 2>&1; echo "pytest_rc=$?"; tail -3 gpurun_out/pytest24.log
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
timeout 300 $CMD > gpurun_out/plain24.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches.csv $CMD > gpurun_out/ncu24.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:voice_render_mix_tma -s 3 -c 1 -f -o gpurun_out/prof_render_c2 $CMD > gpurun_out/ncu24b.log 2>&1; tail -1 gpurun_out/ncu24b.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pcm16_decode_batch -s 3 -c 1 -f -o gpurun_out/prof_decode_r24 $CMD > gpurun_out/ncu24c.log 2>&1; tail -1 gpurun_out/ncu24c.log
