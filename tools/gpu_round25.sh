#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_c1.py > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "c1_rc=$?"; cat gpurun_out/bench_c1.json; tail -3 gpurun_out/bench_c1.err
