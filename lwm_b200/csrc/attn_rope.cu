// Rotary position embedding of the attention prologue (lwm/llama.py:344-375 `precompute_freqs_cis` /
// `apply_rotary_emb`, applied at llama.py:517-519 right before the ring-attention call) — SURVEY.md §8f next-row 2.
//
// The reference gathers rows of a host-built complex64 table [max_pos, D/2] (512 MB at 1 M positions) by
// position_ids and multiplies in fp32 complex arithmetic. Here the table is never materialised: the angle of
// (position, pair j) is rebuilt exactly as the table builder does — float32(float64(pos) * float64(inv_freq[j])),
// np.outer of an int64 and a float32 vector — and cos/sin of that float32 angle are taken in double precision and
// rounded once, i.e. within 0.5 ulp of the true value (numpy's float32 sin/cos are within 1 ulp of it). One CTA
// serves kPos token positions: 256 threads build the kPos*64 (cos, sin) pairs once in shared memory, then every
// thread rotates 8-element vectors of all heads of Q and K for those positions, so the transcendental cost is
// amortised over H heads and the kernel is a pure HBM stream: each element is read once and written once.
// The [B,S,H*D] projection output and the op's [B,S,H,D] input are the same memory: the head split is free.
#include <cuda_bf16.h>

#include "capi_internal.h"

namespace lwm {

constexpr int kRopeDim = 128;
constexpr int kRopePairs = kRopeDim / 2;
constexpr int kRopePos = 4;

// 8 consecutive elements as loaded (kept raw until use, so that a batch of loads costs few registers)
template <typename T>
struct Raw8;
template <>
struct Raw8<float> {
  float4 a, b;
  static constexpr int kBatch = 4;
  __device__ __forceinline__ void load(const float* p) {
    a = reinterpret_cast<const float4*>(p)[0];
    b = reinterpret_cast<const float4*>(p)[1];
  }
  __device__ __forceinline__ void unpack(float (&x)[8]) const {
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  }
};
template <>
struct Raw8<__nv_bfloat16> {
  uint4 a;
  static constexpr int kBatch = 8;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { a = reinterpret_cast<const uint4*>(p)[0]; }
  __device__ __forceinline__ void unpack(float (&x)[8]) const {
    const unsigned w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      x[2 * i] = __uint_as_float(w[i] << 16);
      x[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};
template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&y)[8]);
template <>
__device__ __forceinline__ void store8<float>(float* p, const float (&y)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(y[0], y[1], y[2], y[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(y[4], y[5], y[6], y[7]);
}
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float (&y)[8]) {
  uint4 o;
  unsigned* w = reinterpret_cast<unsigned*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(y[2 * i], y[2 * i + 1]);
    w[i] = *reinterpret_cast<const unsigned*>(&v);
  }
  reinterpret_cast<uint4*>(p)[0] = o;
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256, 3) rope_kernel(const TIn* __restrict__ xq, const TIn* __restrict__ xk,
                                                   TOut* __restrict__ oq, TOut* __restrict__ ok,
                                                   const int* __restrict__ position_ids,
                                                   const float* __restrict__ inv_freq, long long n_tok, int Hq,
                                                   int Hk, float sin_sign) {
  __shared__ float2 cs[kRopePos][kRopePairs];
  const long long tok0 = (long long)blockIdx.x * kRopePos;
  {
    const int p = threadIdx.x >> 6, j = threadIdx.x & 63;
    const long long tok = tok0 + p;
    if (tok < n_tok) {
      const float angle = (float)((double)position_ids[tok] * (double)inv_freq[j]);
      double s, c;
      sincos((double)angle, &s, &c);
      cs[p][j] = make_float2((float)c, sin_sign * (float)s);
    }
  }
  __syncthreads();
  const int vq = Hq * (kRopeDim / 8), vk = Hk * (kRopeDim / 8), vt = vq + vk;
  // batches of independent 16/32-byte loads per thread (8 / 4) before the first use: ~100 KB in flight per SM
  constexpr int kBatch = Raw8<TIn>::kBatch;
  const int total = kRopePos * vt;
  for (int base = threadIdx.x; base < total; base += kBatch * blockDim.x) {
    Raw8<TIn> raw[kBatch];
    int off[kBatch];   // element offset from this CTA's first token (q or k tensor)
    int cs_idx[kBatch];
    bool live[kBatch], is_q[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int v = base + u * blockDim.x;
      const int p = v / vt, r = v - p * vt;
      const long long tok = tok0 + p;
      live[u] = v < total && tok < n_tok;
      is_q[u] = r < vq;
      const int rr = is_q[u] ? r : r - vq;
      off[u] = (p * (is_q[u] ? Hq : Hk)) * kRopeDim + rr * 8;
      cs_idx[u] = p * kRopePairs + (rr & 15) * 4;
      if (live[u]) raw[u].load((is_q[u] ? xq + tok0 * Hq * kRopeDim : xk + tok0 * Hk * kRopeDim) + off[u]);
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      if (!live[u]) continue;
      float x[8], y[8];
      raw[u].unpack(x);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = (&cs[0][0])[cs_idx[u] + i];
        // (a + ib)(c + is) = (ac - bs) + i(as + bc), separately rounded products as in a plain complex64 multiply
        y[2 * i] = __fsub_rn(__fmul_rn(x[2 * i], f.x), __fmul_rn(x[2 * i + 1], f.y));
        y[2 * i + 1] = __fadd_rn(__fmul_rn(x[2 * i], f.y), __fmul_rn(x[2 * i + 1], f.x));
      }
      store8<TOut>((is_q[u] ? oq + tok0 * Hq * kRopeDim : ok + tok0 * Hk * kRopeDim) + off[u], y);
    }
  }
}

template <typename TIn, typename TOut>
static int launch_rope(const void* xq, const void* xk, void* oq, void* ok, const int* pos, const float* inv_freq,
                       long long n_tok, int Hq, int Hk, int conj, cudaStream_t stream) {
  const long long blocks = (n_tok + kRopePos - 1) / kRopePos;
  rope_kernel<TIn, TOut><<<unsigned(blocks), 256, 0, stream>>>(
      reinterpret_cast<const TIn*>(xq), reinterpret_cast<const TIn*>(xk), reinterpret_cast<TOut*>(oq),
      reinterpret_cast<TOut*>(ok), pos, inv_freq, n_tok, Hq, Hk, conj ? -1.0f : 1.0f);
  return lwm_check_launch("rope_kernel");
}

}  // namespace lwm

using namespace lwm;

extern "C" int lwm_attn_rope(const void* xq, const void* xk, int in_dtype, void* out_q, void* out_k, int out_dtype,
                             const int* position_ids, const float* inv_freq, int B, int S, int Hq, int Hk, int D,
                             int conj, void* stream) {
  if (D != kRopeDim) return lwm_fail(LWM_ERR_SHAPE, "attn_rope: head_dim must be 128");
  if (B <= 0 || S <= 0 || Hq <= 0 || Hk < 0) return lwm_fail(LWM_ERR_SHAPE, "attn_rope: bad sizes");
  if (!xq || !out_q || !position_ids || !inv_freq || (Hk > 0 && (!xk || !out_k)))
    return lwm_fail(LWM_ERR_ARG, "attn_rope: null pointer");
  if ((in_dtype != 0 && in_dtype != 1) || (out_dtype != 0 && out_dtype != 1))
    return lwm_fail(LWM_ERR_ARG, "attn_rope: dtype codes are 0 (fp32) or 1 (bf16)");
  if (!lwm_check_device()) return LWM_ERR_DEVICE;
  const long long n_tok = (long long)B * S;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (in_dtype == 0 && out_dtype == 0)
    return launch_rope<float, float>(xq, xk, out_q, out_k, position_ids, inv_freq, n_tok, Hq, Hk, conj, st);
  if (in_dtype == 0 && out_dtype == 1)
    return launch_rope<float, __nv_bfloat16>(xq, xk, out_q, out_k, position_ids, inv_freq, n_tok, Hq, Hk, conj, st);
  if (in_dtype == 1 && out_dtype == 0)
    return launch_rope<__nv_bfloat16, float>(xq, xk, out_q, out_k, position_ids, inv_freq, n_tok, Hq, Hk, conj, st);
  return launch_rope<__nv_bfloat16, __nv_bfloat16>(xq, xk, out_q, out_k, position_ids, inv_freq, n_tok, Hq, Hk, conj, st);
}
