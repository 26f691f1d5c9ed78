// Microbenchmark of tcgen05.mma issue/execute rates on one SM (per-CTA numbers; grid = #SMs so that
// clocks/power are representative). Reports cycles per MMA instruction for:
//   SS / TS operand sourcing, N = 64/128/256, K-major vs MN-major B, same vs alternating accumulator.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bench_umma tools/bench_umma.cu
#include <cstdio>
#include <vector>
#include "../lwm_b200/csrc/ptx.cuh"
using namespace lwm;

struct Case { int ts; int N; int b_mn; int alt; int a_mn; };

__global__ void __launch_bounds__(128, 1) bench_kernel(Case c, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  // operands: garbage-but-finite bf16 (zeros) — timing only
  for (int i = threadIdx.x; i < (32768 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0) {
    tmem_alloc<512>(&tmem_base_s);
    if (lane_id() == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 32768;
    const uint64_t ad = c.a_mn ? desc_mnmajor_sw128(a0, 16384) : desc_kmajor_sw128(a0);
    const uint64_t bd = c.b_mn ? desc_mnmajor_sw128(b0, 16384) : desc_kmajor_sw128(b0);
    const uint32_t idesc = make_idesc_bf16(128, c.N, c.a_mn, c.b_mn);
    long long t0 = clock64();
    if (leader) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t d = tmem + ((c.alt && (ks & 1)) ? 256 : 0);
          const uint32_t aoff = c.a_mn ? ks * 2048 : ((ks >> 2) * 16384 + (ks & 3) * 32);
          const uint32_t boff = c.b_mn ? ks * 2048 : ((ks >> 2) * 16384 + (ks & 3) * 32);
          if (c.ts) umma_ts(d, tmem + 448 + ks * 8, desc_advance(bd, boff), idesc, 1);
          else umma_ss(d, desc_advance(ad, aoff), desc_advance(bd, boff), idesc, 1);
        }
      }
      umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (lane_id() == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  cudaSetDevice(0);
  long long* d_out;
  cudaMalloc(&d_out, 8);
  const int smem_bytes = 32768 + 65536 + 1024;
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  Case cases[] = {{0, 128, 0, 0, 0}, {0, 128, 0, 1, 0}, {1, 128, 0, 0, 0}, {1, 128, 1, 0, 0}, {0, 128, 1, 0, 0},
                  {0, 128, 1, 0, 1}, {0, 256, 0, 0, 0}, {1, 256, 0, 0, 0}, {0, 64, 0, 0, 0}, {1, 64, 0, 0, 0},
                  {1, 64, 1, 0, 0}, {0, 64, 1, 0, 0}, {1, 128, 1, 1, 0}};
  const int iters = 2000;
  for (int grid : {1, 148}) {
    printf("grid=%d\n", grid);
    for (const Case& c : cases) {
      bench_kernel<<<grid, 128, smem_bytes>>>(c, iters, d_out);
      bench_kernel<<<grid, 128, smem_bytes>>>(c, iters, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc = 0;
      cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
      const double per = double(cyc) / (iters * 8.0);
      printf("  %s N=%3d A:%s B:%s %s : %7.1f cyc/MMA  (%.0f%% of the %d-cycle floor)  %s\n", c.ts ? "TS" : "SS", c.N,
             c.ts ? "tmem " : (c.a_mn ? "MN   " : "K    "), c.b_mn ? "MN" : "K ", c.alt ? "alt-D " : "same-D", per,
             100.0 * (c.N / 2.0) / per, c.N / 2, cudaGetErrorString(e));
    }
  }
  return 0;
}
