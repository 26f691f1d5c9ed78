#!/usr/bin/env bash
# 1-GPU profiling session (ncu): launch list of the bench command, one --set full capture of each attention tile kernel
# at the headline shape (default precision mode), launch list + DRAM bytes of one VQGAN encode. Raw files under
# gpurun_out/, summaries are made from them with tools/ncu_summary.py and copied to profiles/.
set -u
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_bench_seq16384.csv \
  python bench.py --seq 16384 --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/ncu_bench_r02.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 1 -c 1 -o gpurun_out/prof_attn_bwd_128k_r02 -f \
  python tools/perf_attn.py bwd 131072 > gpurun_out/ncu_attn_bwd_r02.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 2 -c 1 -o gpurun_out/prof_attn_fwd_128k_r02 -f \
  python tools/perf_attn.py fwd 131072 > gpurun_out/ncu_attn_fwd_r02.log 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --csv --log-file gpurun_out/launches_r02_vqgan_encode16_fp16x2.csv python tools/prof_vqgan_once.py 16 fp16x2 > gpurun_out/ncu_vq_r02.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_umma -s 1 -c 1 -o gpurun_out/prof_conv_umma_r02 -f \
  python tools/prof_vqgan_once.py 16 fp16x2 > gpurun_out/ncu_conv_r02.log 2>&1
ls -la gpurun_out/*r02*
