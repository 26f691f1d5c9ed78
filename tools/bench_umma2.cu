// Microbenchmark of tcgen05.mma.cta_group::2 (CTA pair, UMMA M = 256) issue/execute rates — the question DESIGN.md
// §3.3b leaves open for the next round: does pairing two SMs remove the ~33 serial cycles an smem-sourced A operand
// costs per instruction with cta_group::1 (SS N=128: 107 cycles vs the 64-cycle floor)?
// NOT YET RUN (written after the round's GPU budget was spent). Run under `timeout 60` — an MMA/commit mistake hangs.
//   cluster (2,1,1) per pair, 74 pairs = 148 SMs; the leader CTA's elected thread issues `iters` x 8 MMAs
//   (K = 16 each, one 128-deep k-loop) into the pair's TMEM, commits with a multicast arrive, both CTAs wait.
// Cases: SS / TS, N = 128 / 256 (B is split across the pair: each CTA holds N/2 rows of B; A: each CTA its own 128 rows).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bench_umma2 tools/bench_umma2.cu
#include <cstdio>
#include <cooperative_groups.h>
#include "../lwm_b200/csrc/ptx.cuh"
using namespace lwm;
namespace cg = cooperative_groups;

struct Case { int ts; int N; };

__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result) {   // whole warp, in BOTH CTAs of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(smem_result)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void umma2_ss(uint32_t d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a_desc), "l"(b_desc),
               "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma2_ts(uint32_t d, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(b_desc),
               "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma2_commit_multicast(uint64_t* bar) {   // arrives on `bar` in both CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)), "h"((uint16_t)0x3) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) bench_kernel(Case c, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  cg::cluster_group cluster = cg::this_cluster();
  const int warp = threadIdx.x >> 5;
  const bool leader_cta = cluster.block_rank() == 0;
  for (int i = threadIdx.x; i < (32768 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0) {
    tmem_alloc2(&tmem_base_s);
    if (lane_id() == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  cluster.sync();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 1) {
    const bool leader_thread = elect_one();
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 32768;
    const uint64_t ad = desc_kmajor_sw128(a0);      // this CTA's 128 rows of A (K-major, 128 deep)
    const uint64_t bd = desc_kmajor_sw128(b0);      // this CTA's N/2 rows of B
    const uint32_t idesc = make_idesc_bf16(256, c.N, false, false);
    long long t0 = clock64();
    if (leader_cta && leader_thread) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t off = (ks >> 2) * 16384 + (ks & 3) * 32;
          if (c.ts) umma2_ts(tmem, tmem + 448 + ks * 8, desc_advance(bd, off), idesc, 1);
          else umma2_ss(tmem, desc_advance(ad, off), desc_advance(bd, off), idesc, 1);
        }
      }
      umma2_commit_multicast(&bar);
    }
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (leader_cta && lane_id() == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  cluster.sync();
  if (warp == 0) tmem_dealloc2(tmem);
}

int main() {
  cudaSetDevice(0);
  long long* d_out;
  cudaMalloc(&d_out, 8);
  const int smem_bytes = 32768 + 65536 + 1024;
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  Case cases[] = {{0, 128}, {1, 128}, {0, 256}, {1, 256}};
  const int iters = 2000;
  for (int grid : {2, 148}) {
    printf("grid=%d CTAs (%d pairs)\n", grid, grid / 2);
    for (const Case& c : cases) {
      bench_kernel<<<grid, 128, smem_bytes>>>(c, iters, d_out);
      bench_kernel<<<grid, 128, smem_bytes>>>(c, iters, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc = 0;
      cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
      const double per = double(cyc) / (iters * 8.0);
      // one M=256 x N x 16 instruction keeps BOTH SMs busy for N/2 cycles at the dense bf16 rate
      printf("  cta_group::2 %s M=256 N=%3d : %7.1f cyc/MMA  (%.0f%% of the %d-cycle floor)  %s\n", c.ts ? "TS" : "SS", c.N,
             per, 100.0 * (c.N / 2.0) / per, c.N / 2, cudaGetErrorString(e));
      if (e != cudaSuccess) return 1;
    }
  }
  return 0;
}
