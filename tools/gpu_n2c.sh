#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_final.json 2> gpurun_out/bench_n${N}_final.err; echo "bench_rc=$?"
grep -v "^NCCL" gpurun_out/bench_n${N}_final.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','n_gpus','ms_per_step','kernel_ms','gpu_launches')}); print(d['config']['parallelism']); print(d['e2e']['value'], d['e2e_parse_dropin']['value'])"
tail -3 gpurun_out/bench_n${N}_final.err | grep -v "OMP_NUM\|^\*\*\*" | cut -c1-300
