#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
N=${1:-4}
W="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
L=gpurun_out/ring${N}_r2b.log
: > $L
echo "=== worker dense (peer)" | tee -a $L
timeout 300 $W --master-port 29571 tests/ring_multi_gpu_worker.py 2>&1 | grep -E "RING_MULTI|Error|error|Traceback" | tail -8 | tee -a $L
echo "=== trace" | tee -a $L
timeout 300 $W --master-port 29581 tools/ring_trace_peer.py 131072 2>&1 | grep -vE "^W0|^\*\*|OMP_NUM" > gpurun_out/ring_trace_peer_n${N}.log
grep -E "^rank|summary" gpurun_out/ring_trace_peer_n${N}.log | tee -a $L
echo "=== bench (peer)" | tee -a $L
timeout 400 $W --master-port 29574 bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline --no-vqgan 2>&1 | tail -1 | tee gpurun_out/bench_n${N}_peer_r2b.json | cut -c1-300 | tee -a $L
