#!/bin/bash
# N-GPU checks (N = the GPUs of the box):   gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_multi.sh'
#   the multi-GPU parity tests (group over real peer memory, torchrun + CUDA IPC), the exchange timings, and bench.py
#   (strong scaling, bus_check inside) with the fused peer-memory reduction and with the NCCL all-reduce.
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_peer_bus_gpu.py -x -q -m gpu > gpurun_out/pytest_multi_n$N.log 2>&1; echo "pytest_rc=$?"; tail -5 gpurun_out/pytest_multi_n$N.log
timeout 300 $RUN --master-port 29531 tests/checks/peer_bus_check.py > gpurun_out/peer_n$N.json 2> gpurun_out/peer_n$N.err; echo "peer_rc=$?"; tail -1 gpurun_out/peer_n$N.json
for mode in p2p p2p2 nccl; do
  timeout 900 $RUN --master-port 29532 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --reduce $mode > gpurun_out/bench_n${N}_$mode.json 2> gpurun_out/bench_n${N}_$mode.err; echo "bench_${mode}_rc=$?"
  grep "^{" gpurun_out/bench_n${N}_$mode.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','kernel_ms','gpu_launches','bus_check','weak')}); print({k:(v.get('ms'),v.get('reduce_ms'),v.get('frac'),v.get('check')) for k,v in (d.get('configs') or {}).items()})"
done
