// Small HBM-bound helpers around the attention tile kernels:
//   lwm_attn_bwd_prep     delta[b,h,s] = sum_d dout*out  (the rowsum(g∘out) term of the reference's
//                         custom_vjp bwd, SURVEY.md Appendix A `bwd`)
//   lwm_cast_f32_to_bf16  final cast of the fp32 gradient accumulators to the input dtype
#include "attn_common.cuh"
#include "capi_internal.h"

namespace lwm {

// one warp per (b, s, h) row of 128 elements: 4 bf16 per lane from each tensor (8-byte loads)
__global__ void bwd_prep_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                                float* __restrict__ delta, int B, int H, int S) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // over B*S*H
  const long long n_rows = (long long)B * S * H;
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31;
  const uint2 a = reinterpret_cast<const uint2*>(out + row * kHeadDim)[lane];
  const uint2 g = reinterpret_cast<const uint2*>(dout + row * kHeadDim)[lane];
  const __nv_bfloat162 a0 = *reinterpret_cast<const __nv_bfloat162*>(&a.x);
  const __nv_bfloat162 a1 = *reinterpret_cast<const __nv_bfloat162*>(&a.y);
  const __nv_bfloat162 g0 = *reinterpret_cast<const __nv_bfloat162*>(&g.x);
  const __nv_bfloat162 g1 = *reinterpret_cast<const __nv_bfloat162*>(&g.y);
  float acc = __low2float(a0) * __low2float(g0) + __high2float(a0) * __high2float(g0) +
              __low2float(a1) * __low2float(g1) + __high2float(a1) * __high2float(g1);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    const int h = int(row % H);
    const long long bs = row / H;
    const int s = int(bs % S);
    const int b = int(bs / S);
    delta[((long long)b * H + h) * S + s] = acc;
  }
}

// same with an fp32 `out` (fp16 precision mode keeps the un-rounded output as residual)
__global__ void bwd_prep_f32_kernel(const float* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                                    float* __restrict__ delta, int B, int H, int S) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long n_rows = (long long)B * S * H;
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31;
  const float4 a = reinterpret_cast<const float4*>(out + row * kHeadDim)[lane];
  const uint2 g = reinterpret_cast<const uint2*>(dout + row * kHeadDim)[lane];
  const __nv_bfloat162 g0 = *reinterpret_cast<const __nv_bfloat162*>(&g.x);
  const __nv_bfloat162 g1 = *reinterpret_cast<const __nv_bfloat162*>(&g.y);
  float acc = a.x * __low2float(g0) + a.y * __high2float(g0) + a.z * __low2float(g1) + a.w * __high2float(g1);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    const int h = int(row % H);
    const long long bs = row / H;
    delta[((long long)(bs / S) * H + h) * S + (bs % S)] = acc;
  }
}

// nlse2 = -lse * log2(e), with rows whose lse sits at the masked level (never saw an unmasked key: padded rows)
// mapped to -inf so that the backward gives them p = 0 — hoisted out of the tile kernel, where every key-tile CTA
// would redo it for every query column.
__global__ void lse_to_nlse2_kernel(const float* __restrict__ lse, float* __restrict__ nlse2, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float l = lse[i];
    nlse2[i] = (l < -1.0e29f) ? -INFINITY : -l * kLog2e;
  }
}

__global__ void cast_f32_bf16_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 f = src[i];
    dst[i] = make_uint2(pack_bf16x2(f.x, f.y), pack_bf16x2(f.z, f.w));
  }
}

__global__ void add_f32_kernel(float4* __restrict__ dst, const float4* __restrict__ src, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 a = dst[i];
    const float4 b = src[i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    dst[i] = a;
  }
}

// |x| max over a bf16 tensor as raw bits (non-negative floats order like unsigned ints)
__global__ void absmax_bf16_kernel(const uint4* __restrict__ x, long long n8, unsigned* __restrict__ out_bits) {
  unsigned m = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const uint4 v = x[i];
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      m = max(m, (w[k] & 0x7fffu) << 16);          // low bf16, sign cleared, as fp32 bits
      m = max(m, w[k] & 0x7fff0000u);              // high bf16
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out_bits, m);
}

// x16 = fp16(x / scale) with scale = 2^(e-12), e = exponent of the tensor's |max| (scale 1 for an all-zero
// tensor): the largest magnitude lands in [2^12, 2^13), so anything down to 2^-26 of it stays a normal fp16
// and the conversion of every such bf16 value (8 significant bits) is exact.
__global__ void bf16_to_scaled_f16_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, long long n8,
                                          const unsigned* __restrict__ absmax_bits, float* __restrict__ scale_out) {
  const unsigned bits = *absmax_bits;
  const int e = int(bits >> 23) - 127;
  const float inv = bits ? __uint_as_float(unsigned(127 - (e - 12)) << 23) : 1.0f;   // 2^(12-e)
  if (blockIdx.x == 0 && threadIdx.x == 0) *scale_out = bits ? __uint_as_float(unsigned(127 + (e - 12)) << 23) : 1.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const uint4 v = x[i];
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
    unsigned o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float lo = __uint_as_float(w[k] << 16) * inv, hi = __uint_as_float(w[k] & 0xffff0000u) * inv;
      o[k] = pack_f16x2(lo, hi);
    }
    y[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

}  // namespace lwm

using namespace lwm;

extern "C" int lwm_attn_bwd_prep(const void* out, const void* dout, float* delta, int B, int H, int Sq, int D,
                                 void* stream) {
  if (D != kHeadDim) return lwm_fail(LWM_ERR_SHAPE, "attn_bwd_prep: head_dim must be 128");
  if (!out || !dout || !delta) return lwm_fail(LWM_ERR_ARG, "attn_bwd_prep: null pointer");
  if (!lwm_check_device()) return LWM_ERR_DEVICE;
  const long long rows = (long long)B * Sq * H;
  const int warps = 8;
  bwd_prep_kernel<<<unsigned((rows + warps - 1) / warps), warps * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(out), reinterpret_cast<const __nv_bfloat16*>(dout), delta, B, H, Sq);
  return lwm_check_launch("bwd_prep_kernel");
}

extern "C" int lwm_attn_bwd_prep_f32(const float* out_f32, const void* dout, float* delta, int B, int H, int Sq, int D,
                                     void* stream) {
  if (D != kHeadDim) return lwm_fail(LWM_ERR_SHAPE, "attn_bwd_prep_f32: head_dim must be 128");
  if (!out_f32 || !dout || !delta) return lwm_fail(LWM_ERR_ARG, "attn_bwd_prep_f32: null pointer");
  if (!lwm_check_device()) return LWM_ERR_DEVICE;
  const long long rows = (long long)B * Sq * H;
  const int warps = 8;
  bwd_prep_f32_kernel<<<unsigned((rows + warps - 1) / warps), warps * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      out_f32, reinterpret_cast<const __nv_bfloat16*>(dout), delta, B, H, Sq);
  return lwm_check_launch("bwd_prep_f32_kernel");
}

extern "C" int lwm_attn_bwd_lse(const float* lse, float* nlse2, long long n, void* stream) {
  if (!lse || !nlse2 || n <= 0) return lwm_fail(LWM_ERR_ARG, "attn_bwd_lse: bad args");
  if (!lwm_check_device()) return LWM_ERR_DEVICE;
  const long long want = (n + 255) / 256;
  lse_to_nlse2_kernel<<<unsigned(want < 148LL * 8 ? want : 148LL * 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      lse, nlse2, n);
  return lwm_check_launch("lse_to_nlse2_kernel");
}

extern "C" int lwm_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream) {
  if (n % 4) return lwm_fail(LWM_ERR_SHAPE, "cast_f32_to_bf16: n must be a multiple of 4");
  if (!lwm_check_device()) return LWM_ERR_DEVICE;
  if (n == 0) return LWM_OK;
  const long long n4 = n / 4;
  const int threads = 256;
  const long long want = (n4 + threads - 1) / threads;
  const unsigned blocks = unsigned(want < 148LL * 16 ? want : 148LL * 16);
  cast_f32_bf16_kernel<<<blocks, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(src), reinterpret_cast<uint2*>(dst), n4);
  return lwm_check_launch("cast_f32_bf16_kernel");
}

extern "C" int lwm_add_f32(float* dst, const float* src, long long n, void* stream) {
  if (n % 4) return lwm_fail(LWM_ERR_SHAPE, "add_f32: n must be a multiple of 4");
  if (!lwm_check_device()) return LWM_ERR_DEVICE;
  if (n == 0) return LWM_OK;
  const long long n4 = n / 4;
  const int threads = 256;
  const long long want = (n4 + threads - 1) / threads;
  const unsigned blocks = unsigned(want < 148LL * 16 ? want : 148LL * 16);
  add_f32_kernel<<<blocks, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<float4*>(dst), reinterpret_cast<const float4*>(src), n4);
  return lwm_check_launch("add_f32_kernel");
}

// bf16 tensor -> exact scaled fp16 copy + its power-of-two scale (device float). workspace: 4 bytes.
extern "C" int lwm_attn_to_f16(const void* src_bf16, void* dst_f16, float* scale_out, void* workspace, long long n,
                               void* stream) {
  if (!src_bf16 || !dst_f16 || !scale_out || !workspace) return lwm_fail(LWM_ERR_ARG, "attn_to_f16: null pointer");
  if (n <= 0 || n % 8) return lwm_fail(LWM_ERR_SHAPE, "attn_to_f16: n must be a positive multiple of 8");
  if (!lwm_check_device()) return LWM_ERR_DEVICE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(workspace, 0, 4, st) != cudaSuccess) return lwm_fail(LWM_ERR_CUDA, "attn_to_f16: memset failed");
  const long long n8 = n / 8;
  const long long want = (n8 + 255) / 256;
  const unsigned blocks = unsigned(want < 148LL * 8 ? want : 148LL * 8);
  absmax_bf16_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(src_bf16), n8,
                                             reinterpret_cast<unsigned*>(workspace));
  bf16_to_scaled_f16_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(src_bf16),
                                                    reinterpret_cast<uint4*>(dst_f16), n8,
                                                    reinterpret_cast<const unsigned*>(workspace), scale_out);
  return lwm_check_launch("attn_to_f16 kernels");
}
