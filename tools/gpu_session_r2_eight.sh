#!/usr/bin/env bash
# N-GPU validation of the peer-memory executor at BASELINE configs[2] (128K tokens over 8 GPUs): dense-oracle worker,
# sampled-oracle worker at the full length, bench (parity + timing). Logs under gpurun_out/.
set -u
mkdir -p gpurun_out
N=${1:-8}
W="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
L=gpurun_out/ring${N}_r2.log
: > $L
echo "=== worker dense (peer)" | tee -a $L
timeout 300 $W --master-port 29571 tests/ring_multi_gpu_worker.py 2>&1 | grep -vE "^W0|^\*\*|OMP_NUM" | grep -E "RING_MULTI|rank 0 |rank $((N-1)) |Error|error" | tail -30 | tee -a $L
echo "=== worker sampled (peer)" | tee -a $L
RING_TEST_MODE=sampled RING_TEST_S=${2:-131072} timeout 300 $W --master-port 29572 tests/ring_multi_gpu_worker.py 2>&1 | grep -vE "^W0|^\*\*|OMP_NUM" | tail -12 | tee -a $L
echo "=== bench (peer)" | tee -a $L
timeout 400 $W --master-port 29574 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-vqgan 2>&1 | tail -1 | tee gpurun_out/bench_n${N}_peer_r2.json | cut -c1-300 | tee -a $L
if [ "${3:-}" = "nccl" ]; then
echo "=== bench (nccl)" | tee -a $L
LWM_RING_TRANSPORT=nccl timeout 400 $W --master-port 29575 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-vqgan --no-parity 2>&1 | tail -1 | tee gpurun_out/bench_n${N}_nccl_r2.json | cut -c1-300 | tee -a $L
fi
