#!/bin/bash
# first GPU pass: parity tests, bench (both layouts), launch list, one full ncu capture of the decode kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> gpurun_out/host.txt; free -g | head -2 >> gpurun_out/host.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest_rc=$?" | tee -a gpurun_out/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke_rc=$?" | tee -a gpurun_out/smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_image.json 2> gpurun_out/bench_image.err; echo "bench_rc=$?"
python bench.py --steps 20 --warmup 3 --layout payload --no-e2e --no-cpu > gpurun_out/bench_payload.json 2> gpurun_out/bench_payload.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pcm16_decode_batch -s 3 -c 2 -o gpurun_out/prof_decode $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/pytest.log; cat gpurun_out/bench_image.json; cat gpurun_out/bench_payload.json
