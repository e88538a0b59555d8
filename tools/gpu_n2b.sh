#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/peer_bus_check.py > gpurun_out/peer_n$N.json 2> gpurun_out/peer_n$N.err; echo "peer_rc=$?"; tail -1 gpurun_out/peer_n$N.json; tail -5 gpurun_out/peer_n$N.err | grep -v "OMP_NUM\|^\*\*\*" | cut -c1-300
for mode in "p2p --peer-mode root" "p2p --peer-mode scatter" nccl; do
tag=$(echo $mode | tr -d ' -')
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-e2e --reduce $mode > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.err; echo "bench_${tag}_rc=$?"
grep -v "^NCCL" gpurun_out/bench_n${N}_$tag.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','n_gpus','ms_per_step','kernel_ms','gpu_launches')})"
tail -2 gpurun_out/bench_n${N}_$tag.err | grep -v "OMP_NUM\|^\*\*\*" | cut -c1-300
done
