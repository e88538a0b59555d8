#!/bin/bash
# The round's N-GPU evidence on one box (N = all its GPUs):   gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_multi8.sh'
#   multi-GPU parity tests, bench.py at N and N/2 (strong scaling, bus_check inside; the three reduction variants), the
#   concurrent-ingest PCIe probe behind the e2e numbers, the sharded MPEG index.
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
summ() { grep "^{" $1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','kernel_ms','bus_check','weak')}); print('e2e',(d.get('e2e') or {}).get('value'), {k:(v.get('ms'),v.get('reduce_ms'),v.get('frac'),v.get('check')) for k,v in (d.get('configs') or {}).items()})"; }
timeout 900 python -m pytest tests/test_peer_bus_gpu.py -x -q -m gpu > gpurun_out/pytest_multi_n$N.log 2>&1; echo "pytest_rc=$?"; tail -3 gpurun_out/pytest_multi_n$N.log
timeout 600 $RUN --nproc-per-node $N --master-port 29532 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}_p2p2.json 2> gpurun_out/bench_n${N}_p2p2.err; echo "bench_rc=$?"; summ gpurun_out/bench_n${N}_p2p2.json
for mode in nccl p2p; do
  timeout 300 $RUN --nproc-per-node $N --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --no-e2e --reduce $mode > gpurun_out/bench_n${N}_$mode.json 2> gpurun_out/bench_n${N}_$mode.err; echo "bench_${mode}_rc=$?"; summ gpurun_out/bench_n${N}_$mode.json
done
H=$((N/2))
timeout 300 $RUN --nproc-per-node $H --master-port 29534 bench.py --gpus $H --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_n${H}_p2p2.json 2> gpurun_out/bench_n${H}_p2p2.err; echo "bench_n${H}_rc=$?"; summ gpurun_out/bench_n${H}_p2p2.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_n$N.json 2> gpurun_out/bench_reference_n$N.err
timeout 300 $RUN --nproc-per-node $N --master-port 29535 tools/pcie_probe_multi.py > gpurun_out/pcie_multi_n$N.json 2> gpurun_out/pcie_multi_n$N.err; echo "pcie_rc=$?"; python -c "
import json; d=json.loads([l for l in open('gpurun_out/pcie_multi_n$N.json') if l.startswith('{')][-1]); [print(r['concurrent_ranks'], r['h2d_sum_GBps'], r['d2h_sum_GBps'], r['bidir_each_sum_GBps']) for r in d['rows']]"
timeout 400 $RUN --nproc-per-node $N --master-port 29536 tests/checks/mpeg_sharded_check.py --gib 4 > gpurun_out/mpeg_n$N.json 2> gpurun_out/mpeg_n$N.err; echo "mpeg_rc=$?"; tail -1 gpurun_out/mpeg_n$N.json | cut -c1-300
timeout 200 $RUN --nproc-per-node $N --master-port 29537 tests/checks/peer_bus_check.py > gpurun_out/peer_n$N.json 2> gpurun_out/peer_n$N.err; echo "peer_rc=$?"; tail -1 gpurun_out/peer_n$N.json
