// Building-block probe for the sm_100a primitives the kernels in lwm_b200/csrc rely on.
// Each case runs one 128x128x128 bf16 UMMA with a different operand source/layout and checks
// the fp32 result (read back with tcgen05.ld) against a host loop. Small-integer data => exact.
//
//   case  A operand                    B operand
//   0     smem, K-major  (TMA SW128)   smem, K-major          D = A  * B^T   (Q K^T)
//   1     smem, K-major                smem, MN-major         D = A  * Bm    (P V with P in smem)
//   2     TMEM (tcgen05.st packed)     smem, MN-major         D = A  * Bm    (P V with P in TMEM)
//   3     smem, MN-major               smem, MN-major         D = Am^T * Bm  (dQ = dS K)
//   4     TMEM                         smem, K-major          D = A  * B^T
// plus: TMA fp32 reduce-add of a 128x64 tile, and a 4D TMA load with negative / OOB coords.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/probe_umma tools/probe_umma.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_fp16.h>
#include "../lwm_b200/csrc/ptx.cuh"
#include "../lwm_b200/csrc/tmap.h"

using namespace lwm;

struct ProbeParams {
  int a_mode;  // 0 smem K-major, 1 smem MN-major, 2 TMEM
  int b_mode;  // 0 smem K-major, 1 smem MN-major
  int a_f16;   // 1: the A operand holds IEEE fp16 while B stays bf16 (mixed-format kind::f16 MMA)
};

__global__ void __launch_bounds__(160, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __nv_bfloat16* __restrict__ a_rowmajor, float* __restrict__ d_out, ProbeParams pp) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;           // 32 KB: two 16 KB chunks
  uint8_t* sB = smem + 32768;   // 32 KB
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5;
  if (warp == 4) {
    tmem_alloc<512>(&tmem_base_s);
    if (lane_id() == 0) {
      mbar_init(&bar_load, 1);
      mbar_init(&bar_mma, 1);
      fence_mbar_init();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 4 && lane_id() == 0) {
    uint32_t bytes = 32768 + (pp.a_mode == 2 ? 0 : 32768);
    mbar_arrive_expect_tx(&bar_load, bytes);
    if (pp.a_mode != 2) {
      tma_load_2d(sA, &tmA, &bar_load, 0, 0);
      tma_load_2d(sA + 16384, &tmA, &bar_load, 64, 0);
    }
    tma_load_2d(sB, &tmB, &bar_load, 0, 0);
    tma_load_2d(sB + 16384, &tmB, &bar_load, 64, 0);
  }
  if (warp < 4 && pp.a_mode == 2) {
    // write A (row = this thread's lane) into TMEM columns [256, 320) as packed bf16 pairs
    const int row = threadIdx.x;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a_rowmajor + row * 128);
    uint32_t v[32];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = src[h * 32 + i];
      tmem_st_x32(tmem + (uint32_t(warp * 32) << 16) + 256 + h * 32, v);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 4 && lane_id() == 0) {
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc(128, 128, pp.a_mode == 1, pp.b_mode == 1, pp.a_f16 ? kFmtF16 : kFmtBF16, kFmtBF16);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      uint64_t bd = (pp.b_mode == 0) ? desc_kmajor_sw128(b0 + (ks >> 2) * 16384 + (ks & 3) * 32)
                                     : desc_mnmajor_sw128(b0 + ks * 2048, 16384);
      if (pp.a_mode == 2) {
        umma_ts(tmem, tmem + 256 + ks * 8, bd, idesc, ks > 0);
      } else {
        uint64_t ad = (pp.a_mode == 0) ? desc_kmajor_sw128(a0 + (ks >> 2) * 16384 + (ks & 3) * 32)
                                       : desc_mnmajor_sw128(a0 + ks * 2048, 16384);
        umma_ss(tmem, ad, bd, idesc, ks > 0);
      }
    }
    umma_commit(&bar_mma);
  }
  if (warp < 4) {
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    const int row = threadIdx.x;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld_x32(tmem + (uint32_t(warp * 32) << 16) + c * 32, v);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) d_out[row * 128 + c * 32 + i] = __uint_as_float(v[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<512>(tmem);
}

// TMA reduce-add probe: smem tile [128 rows][32 fp32] (128 B rows, no swizzle) added into global [128][32]
__global__ void reduce_probe(const __grid_constant__ CUtensorMap tmR) {
  __shared__ __align__(1024) float tile[128 * 32];
  for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) tile[i] = float(i % 7);
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    tma_reduce_add_4d(&tmR, tile, 0, 0, 0, 0);
    tma_commit_group();
    tma_wait_group<0>();
  }
}

// 4D load with OOB: tensor [N=2][H=4][W=4][C=64] bf16, box (64,4,4,1) at (0,-1,-1,1): expect zero halo
__global__ void oob_probe(const __grid_constant__ CUtensorMap tm, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, 16 * 128);
    tma_load_4d(smem, &tm, &bar, 0, -1, -1, 1);
  }
  mbar_wait(&bar, 0);
  // un-swizzle: pixel p (0..15), channel 0
  if (threadIdx.x < 16) {
    int p = threadIdx.x;
    const __nv_bfloat16* rowp = reinterpret_cast<const __nv_bfloat16*>(smem + swz128_offset(p, 0));
    out[p] = __bfloat162float(rowp[0]);
  }
}


int main() {
  cudaSetDevice(0);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  printf("device: %s sm_%d%d SMs=%d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  const int M = 128, N = 128, K = 128;
  std::vector<float> A(M * K), B(N * K);  // logical A[m][k], B[n][k]
  srand(1);
  for (auto& x : A) x = float(rand() % 5 - 2);
  for (auto& x : B) x = float(rand() % 7 - 3);
  std::vector<float> ref(M * N, 0.f);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k];
      ref[m * N + n] = s;
    }
  std::vector<__nv_bfloat16> hA_k(M * K), hA_mn(K * M), hB_k(N * K), hB_mn(K * N);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      hA_k[m * K + k] = __float2bfloat16(A[m * K + k]);
      hA_mn[k * M + m] = __float2bfloat16(A[m * K + k]);
    }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      hB_k[n * K + k] = __float2bfloat16(B[n * K + k]);
      hB_mn[k * N + n] = __float2bfloat16(B[n * K + k]);
    }
  __nv_bfloat16 *dA_k, *dA_mn, *dB_k, *dB_mn;
  float* dD;
  cudaMalloc(&dA_k, M * K * 2); cudaMalloc(&dA_mn, M * K * 2);
  cudaMalloc(&dB_k, N * K * 2); cudaMalloc(&dB_mn, N * K * 2);
  cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA_k, hA_k.data(), M * K * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dA_mn, hA_mn.data(), M * K * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB_k, hB_k.data(), N * K * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB_mn, hB_mn.data(), N * K * 2, cudaMemcpyHostToDevice);

  // all four are [128 rows][128 cols] bf16 row-major; box = 64 cols x 128 rows, SW128
  auto mk = [&](void* p) {
    CUtensorMap t;
    uint64_t dims[2] = {128, 128};
    uint64_t strides[1] = {256};
    uint32_t box[2] = {64, 128};
    if (!lwm::encode_tmap(&t, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p, dims, strides, box,
                          CU_TENSOR_MAP_SWIZZLE_128B)) {
      printf("tensor map encode failed\n");
      exit(2);
    }
    return t;
  };
  CUtensorMap tA_k = mk(dA_k), tA_mn = mk(dA_mn), tB_k = mk(dB_k), tB_mn = mk(dB_mn);

  const int smem_bytes = 65536 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  // fp16 copies of A (same logical values) for the mixed-format cases
  std::vector<__half> hA_k16(M * K), hA_mn16(K * M);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      hA_k16[m * K + k] = __float2half(A[m * K + k]);
      hA_mn16[k * M + m] = __float2half(A[m * K + k]);
    }
  __nv_bfloat16 *dA_k16, *dA_mn16;
  cudaMalloc(&dA_k16, M * K * 2); cudaMalloc(&dA_mn16, M * K * 2);
  cudaMemcpy(dA_k16, hA_k16.data(), M * K * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dA_mn16, hA_mn16.data(), M * K * 2, cudaMemcpyHostToDevice);
  CUtensorMap tA_k16 = mk(dA_k16), tA_mn16 = mk(dA_mn16);
  int cases[9][3] = {{0, 0, 0}, {0, 1, 0}, {2, 1, 0}, {1, 1, 0}, {2, 0, 0}, {2, 1, 1}, {0, 1, 1}, {1, 1, 1}, {0, 0, 1}};
  int fails = 0;
  for (int c = 0; c < 9; ++c) {
    ProbeParams pp{cases[c][0], cases[c][1], cases[c][2]};
    if (pp.a_f16) {
      cudaMemset(dD, 0xff, M * N * 4);
      probe_kernel<<<1, 160, smem_bytes>>>(pp.a_mode == 1 ? tA_mn16 : tA_k16, pp.b_mode == 1 ? tB_mn : tB_k, dA_k16,
                                           dD, pp);
    } else {
    cudaMemset(dD, 0xff, M * N * 4);
    probe_kernel<<<1, 160, smem_bytes>>>(pp.a_mode == 1 ? tA_mn : tA_k, pp.b_mode == 1 ? tB_mn : tB_k, dA_k,
                                         dD, pp);
    }
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> out(M * N);
    cudaMemcpy(out.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    int bad = 0;
    for (int i = 0; i < M * N; ++i) {
      double d = std::fabs(double(out[i]) - ref[i]);
      if (!(d <= 1e-3)) ++bad;
      if (d > maxerr || d != d) maxerr = d;
    }
    printf("case %d (a_mode=%d b_mode=%d a_f16=%d): cuda=%s bad=%d/%d maxerr=%g  D[0][0..3]=%g %g %g %g ref=%g %g %g %g\n",
           c, pp.a_mode, pp.b_mode, pp.a_f16, cudaGetErrorString(e), bad, M * N, maxerr, out[0], out[1], out[2], out[3],
           ref[0], ref[1], ref[2], ref[3]);
    if (bad || e != cudaSuccess) ++fails;
    if (e != cudaSuccess) { printf("sticky error, abort\n"); return 1; }
  }

  {  // reduce-add probe
    float* dR;
    cudaMalloc(&dR, 128 * 32 * 4);
    std::vector<float> h(128 * 32, 1.0f);
    cudaMemcpy(dR, h.data(), 128 * 32 * 4, cudaMemcpyHostToDevice);
    CUtensorMap t;
    uint64_t dims[4] = {32, 128, 1, 1};
    uint64_t strides[3] = {128, 128 * 128, 128 * 128};
    uint32_t box[4] = {32, 128, 1, 1};
    bool ok = lwm::encode_tmap(&t, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dR, dims, strides, box,
                               CU_TENSOR_MAP_SWIZZLE_NONE);
    reduce_probe<<<1, 128>>>(t);
    reduce_probe<<<1, 128>>>(t);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h.data(), dR, 128 * 32 * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < 128 * 32; ++i) bad += (h[i] != 1.0f + 2.0f * float(i % 7));
    printf("reduce-add: encode=%d cuda=%s bad=%d\n", int(ok), cudaGetErrorString(e), bad);
    if (bad || e != cudaSuccess) ++fails;
  }
  {  // OOB probe
    const int NN = 2, H = 4, W = 4, C = 64;
    std::vector<__nv_bfloat16> hx(NN * H * W * C);
    for (int n = 0; n < NN; ++n)
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
          for (int c = 0; c < C; ++c) hx[((n * H + y) * W + x) * C + c] = __float2bfloat16(float(100 * n + 10 * y + x + 1));
    __nv_bfloat16* dx;
    float* dout;
    cudaMalloc(&dx, hx.size() * 2);
    cudaMalloc(&dout, 16 * 4);
    cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap t;
    uint64_t dims[4] = {uint64_t(C), uint64_t(W), uint64_t(H), uint64_t(NN)};
    uint64_t strides[3] = {uint64_t(C) * 2, uint64_t(W) * C * 2, uint64_t(H) * W * C * 2};
    uint32_t box[4] = {64, 4, 4, 1};
    bool ok = lwm::encode_tmap(&t, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dx, dims, strides, box,
                               CU_TENSOR_MAP_SWIZZLE_128B);
    cudaFuncSetAttribute(oob_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 + 1024);
    oob_probe<<<1, 32, 4096 + 1024>>>(t, dout);
    cudaError_t e = cudaDeviceSynchronize();
    float ho[16];
    cudaMemcpy(ho, dout, 64, cudaMemcpyDeviceToHost);
    printf("oob 4d (expect row0/col0 zero, rest 100+10*(y-1)+(x-1)+1): encode=%d cuda=%s\n", int(ok),
           cudaGetErrorString(e));
    int bad = 0;
    for (int y = 0; y < 4; ++y) {
      for (int x = 0; x < 4; ++x) {
        float expect = (y == 0 || x == 0) ? 0.f : float(100 + 10 * (y - 1) + (x - 1) + 1);
        bad += (ho[y * 4 + x] != expect);
        printf(" %6.0f", ho[y * 4 + x]);
      }
      printf("\n");
    }
    if (bad || e != cudaSuccess) ++fails;
  }
  printf("PROBE %s (%d failing)\n", fails ? "FAIL" : "OK", fails);
  return fails ? 1 : 0;
}
