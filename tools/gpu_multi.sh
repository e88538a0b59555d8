#!/bin/bash
# N-GPU checks (N = the GPUs of the box):   gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_multi.sh'
#   the mix reduction over peer memory (both choreographies) vs NCCL vs a single-GPU render, the frame-offset index
#   of one MPEG stream cut into byte ranges vs the oracle, and bench.py in its three reduction variants.
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29531 tests/checks/peer_bus_check.py > gpurun_out/peer_n$N.json 2> gpurun_out/peer_n$N.err; echo "peer_rc=$?"; tail -1 gpurun_out/peer_n$N.json
timeout 600 $RUN --master-port 29533 tests/checks/mpeg_sharded_check.py --gib 8 > gpurun_out/mpeg_n$N.json 2> gpurun_out/mpeg_n$N.err; echo "mpeg_rc=$?"; tail -1 gpurun_out/mpeg_n$N.json
for mode in "p2p --peer-mode root" "p2p --peer-mode scatter" nccl; do
  tag=$(echo $mode | tr -d ' -')
  timeout 600 $RUN --master-port 29532 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --reduce $mode > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.err; echo "bench_${tag}_rc=$?"
  grep -v "^NCCL" gpurun_out/bench_n${N}_$tag.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','n_gpus','ms_per_step','kernel_ms','gpu_launches')})"
done
