#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/pcie_probe.py > gpurun_out/pcie.json 2> gpurun_out/pcie.err; cat gpurun_out/pcie.json
timeout 400 python tools/bench_mpeg.py > gpurun_out/bench_mpeg_v2.json 2> gpurun_out/bench_mpeg_v2.err; echo "bench_mpeg_rc=$?"
cat gpurun_out/bench_mpeg_v2.json; tail -3 gpurun_out/bench_mpeg_v2.err
timeout 600 python bench.py > gpurun_out/bench_r6.json 2> gpurun_out/bench_r6.err; echo "bench_rc=$?"; tail -3 gpurun_out/bench_r6.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r6.json 2> gpurun_out/bench_ref_r6.err; echo "ref_rc=$?"
CMD="python tools/bench_mpeg.py --gib 4 --iters 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mpeg_sync_scan -s 1 -c 1 -o gpurun_out/prof_mpeg_v2 $CMD > gpurun_out/ncu_mpeg_v2.log 2>&1
tail -2 gpurun_out/ncu_mpeg_v2.log
nproc; lscpu | grep "Model name"
