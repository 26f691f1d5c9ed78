#!/usr/bin/env bash
# First GPU calls of the next round: everything here was built and CPU-tested after round 1's GPU budget ran out and has
# NOT run on hardware yet. Logs go to gpurun_out/.
#   gpurun --timeout 900 -- 'bash tools/next_gpu_session.sh one'          # 1 GPU
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/next_gpu_session.sh two'  # 2 GPUs (charged 2x)
set -u
mkdir -p gpurun_out
W="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571"
case "${1:-one}" in
  one)
    timeout 60 tools/bench_umma2 2>&1 | tee gpurun_out/umma2_rates.log           # cta_group::2 issue rates (may hang: timeout)
    timeout 300 python tools/compare_flash_attn.py 4096 16384 131072 2>&1 | tee gpurun_out/compare_flash_attn.log
    timeout 120 python tools/perf_vqgan.py 16 bf16x3 2>&1 | tee gpurun_out/vqgan_perf.log   # now incl. decode
    timeout 300 python bench.py --seq 16384 --steps 3 --warmup 3 --e2e-overlap --no-cpu-baseline --no-vqgan 2>&1 \
      | tail -1 > gpurun_out/bench_e2e_overlap_16k.json
    timeout 300 python bench.py --seq 16384 --steps 3 --warmup 3 --precision fp16 --no-cpu-baseline --no-vqgan 2>&1 \
      | tail -1 > gpurun_out/bench_fp16_16k.json
    ;;
  two)
    for env in "" "LWM_ATTN_PRECISION=fp16" "LWM_RING_PREFETCH=all" "LWM_RING_TRANSPORT=symm" \
               "LWM_RING_TRANSPORT=symm LWM_ATTN_PRECISION=fp16"; do
      echo "=== ${env:-default} ===" | tee -a gpurun_out/ring2_modes.log
      env $env timeout 300 $W tests/ring_multi_gpu_worker.py 2>&1 | tail -12 | tee -a gpurun_out/ring2_modes.log
    done
    for env in "" "LWM_RING_PREFETCH=all" "LWM_RING_TRANSPORT=symm"; do
      echo "=== bench N=2 ${env:-default} ===" | tee -a gpurun_out/ring2_modes.log
      env $env timeout 400 $W bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-vqgan 2>&1 | tail -1 \
        | tee -a gpurun_out/ring2_modes.log
    done
    ;;
esac
