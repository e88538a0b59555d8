#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench_rc=$?"
grep -v "^NCCL" gpurun_out/bench_n$N.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','n_gpus','ms_per_step','kernel_ms','gpu_launches')}); print({k:v for k,v in d['e2e'].items() if k!='api'})"
tail -3 gpurun_out/bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref_rc=$?"; cut -c1-300 gpurun_out/bench_ref_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 tools/mpeg_sharded_check.py --gib 8 > gpurun_out/mpeg_n$N.json 2> gpurun_out/mpeg_n$N.err; echo "mpeg_rc=$?"; tail -1 gpurun_out/mpeg_n$N.json; tail -3 gpurun_out/mpeg_n$N.err
