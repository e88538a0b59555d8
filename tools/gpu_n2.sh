#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mpeg_sharded_check.py --gib 8 > gpurun_out/mpeg_n2.json 2> gpurun_out/mpeg_n2.err; echo "mpeg_n2_rc=$?"; tail -3 gpurun_out/mpeg_n2.json; tail -5 gpurun_out/mpeg_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench_n2_rc=$?"; tail -c 1500 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref_n2_rc=$?"
