#!/usr/bin/env bash
# 2-GPU validation of the peer-memory executor: dense-oracle worker, BASELINE-length sampled-oracle worker, the NCCL
# transport as regression, and the bench on both transports. Logs under gpurun_out/.
set -u
mkdir -p gpurun_out
N=${1:-2}
W="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
L=gpurun_out/ring${N}_r2.log
: > $L
echo "=== worker dense (peer)" | tee -a $L
timeout 240 $W --master-port 29571 tests/ring_multi_gpu_worker.py 2>&1 | grep -vE "^W0|^\*\*|OMP_NUM" | tail -40 | tee -a $L
echo "=== worker sampled (peer)" | tee -a $L
RING_TEST_MODE=sampled RING_TEST_S=${2:-32768} timeout 240 $W --master-port 29572 tests/ring_multi_gpu_worker.py 2>&1 | grep -vE "^W0|^\*\*|OMP_NUM" | tail -12 | tee -a $L
echo "=== worker dense (nccl)" | tee -a $L
LWM_RING_TRANSPORT=nccl timeout 240 $W --master-port 29573 tests/ring_multi_gpu_worker.py 2>&1 | grep -vE "^W0|^\*\*|OMP_NUM" | tail -6 | tee -a $L
echo "=== bench (peer)" | tee -a $L
timeout 400 $W --master-port 29574 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-vqgan 2>&1 | tail -1 | tee gpurun_out/bench_n${N}_peer_r2.json | cut -c1-400 | tee -a $L
echo "=== bench (nccl)" | tee -a $L
LWM_RING_TRANSPORT=nccl timeout 400 $W --master-port 29575 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-vqgan --no-parity 2>&1 | tail -1 | tee gpurun_out/bench_n${N}_nccl_r2.json | cut -c1-400 | tee -a $L
