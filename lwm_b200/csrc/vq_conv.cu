// VQGAN convolutions (flax nn.Conv of lwm/vqgan.py: ResnetBlock 3x3, 1x1 shortcuts, Downsample
// stride-2, Upsample conv, conv_out, quant/post_quant 1x1) as a persistent implicit GEMM on the
// 5th-gen tensor cores.
//
//   GEMM view  M = N*Ho*Wo output pixels, N = Cout, K = taps * Cin.
//   A operand  the activation plane [N,Hin,Win,Cpad] (bf16, written by lwm_vq_prep) is never
//              im2col'ed in memory: for tap (kh,kw) and a 64-channel slice, ONE 4-D TMA box
//              (64 ch x 16 px x 8 rows) lands as a 128-row x 128-byte K-major swizzled tile; the
//              tap shift is just a coordinate offset and out-of-image pixels are zero-filled by
//              the TMA unit (SAME padding / the bottom-right pad of Downsample for free). Stride-2
//              convs use the tensor map's element strides.
//   B operand  weights pre-packed [tap][Cout_pad][Cpad] bf16 (K-major), 3-D TMA box (64, BN, 1).
//   precision  n_pass = 1: bf16 x bf16 (fast). n_pass = 3: both operands split x = hi + lo (two bf16)
//              and D += A_hi B_hi + A_lo B_hi + A_hi B_lo — fp32-class accuracy (2^-16 products,
//              fp32 accumulation in TMEM), which is what the reference's fp32 convs need.
//              n_pass = 2 ("fp16x2"): the activation is ONE fp16 plane and the weights are split w = hi + lo (two fp16,
//              pre-scaled by a power of two); both halves are stacked along N ([BN hi rows | BN lo rows]) so a single
//              128 x 2BN x 16 UMMA produces A.hi and A.lo side by side in TMEM and the epilogue adds the two halves:
//              2x the algorithmic MMA work instead of 3x, wide-N instructions (the measured SS issue rate is 171
//              cycles at N=256 vs 2 x 107 at N=128), one operand plane to write and read instead of two, and
//              8.9e-4 end-to-end relative error on the encoder (activation rounding to 11 bits is all that is left).
//   GN stats   optional epilogue: per-(sample, group) sum / sum of squares of the conv OUTPUT (bias and residual
//              included) — the statistics the next GroupNorm needs — so no separate pass re-reads the activation.
//   pipeline   warp 4: TMA producer; warp 5: UMMA issuer; warps 0-3: epilogue (TMEM -> registers
//              -> + bias (+ residual) -> fp32 NHWC). Accumulators are double-buffered in TMEM so the
//              epilogue of tile i overlaps the main loop of tile i+1. Grid = #SMs, tiles strided.
#include "ptx.cuh"
#include "tmap.h"
#include "capi_internal.h"

namespace lwm {

constexpr int kConvThreads = 192;
constexpr int kATile = 128 * 128;  // 16 KB: 128 pixels x 64 bf16

struct ConvParams {
  int N, Ho, Wo, Cout, Cpad;      // output geometry, real Cout, padded Cin
  int taps_w, taps;               // 3 (or 1), 9 (or 1)
  int stride, pad;                // input coordinate = out*stride + tap - pad
  int BN, n_tiles;                // N-tile width (<= 256, multiple of 16) and count
  int n_pass;                     // 1 (bf16), 3 (bf16x3) or 2 (fp16 activation x stacked fp16 hi|lo weights)
  float w_scale_inv;              // n_pass 2: the weights were packed multiplied by 1/w_scale_inv (a power of two)
  double* stats;                  // optional [N, groups, 2] (sum, sumsq) of the output, accumulated with atomics
  int groups;
  int stages;
  int clip;                       // 1: clamp the result to [-1, 1] (VQGANModel.decode, vqgan.py:141)
  const float* bias;              // [Cout]
  const float* residual;          // [N,Ho,Wo,Cout] or null
  float* out;                     // [N,Ho,Wo,Cout]
};

__global__ void __launch_bounds__(kConvThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
                 const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
                 const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023u) __trap();
  const int b_bytes = (p.n_pass == 2 ? 2 : 1) * p.BN * 128;
  const int stage_bytes = (p.n_pass == 3 ? 2 : 1) * (kATile + b_bytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty = full + 8;
  uint64_t* tmem_full = empty + 8;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_stats = reinterpret_cast<float*>(tmem_base_s + 2);   // [64][2] per-group partials of the current tile

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_w = p.Wo / 16, tiles_h = p.Ho / 8;
  const int m_tiles = p.N * tiles_h * tiles_w;
  const int total_tiles = m_tiles * p.n_tiles;
  const int k_chunks = p.Cpad / 64;
  const int k_iters = p.taps * k_chunks;

  if (warp == 5) {
    tmem_alloc<512>(tmem_base_s);
  } else if (warp == 4 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128);
    }
    fence_mbar_init();
    for (int i = 0; i < 128; ++i) s_stats[i] = 0.f;
    tma_prefetch_desc(&tmAhi);
    tma_prefetch_desc(&tmBhi);
    if (p.n_pass == 3) {
      tma_prefetch_desc(&tmAlo);
      tma_prefetch_desc(&tmBlo);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_base_s;

  // every CTA owns a CONTIGUOUS range of tiles, ordered N-tile-major then image-major: consecutive tiles share the
  // activation halo in L2, and (image, N tile) — the key of the GroupNorm-statistics accumulators — changes at most a
  // few times per CTA
  const int tile_begin = int((long long)total_tiles * blockIdx.x / gridDim.x);
  const int tile_end = int((long long)total_tiles * (blockIdx.x + 1) / gridDim.x);
  auto decode_tile = [&](int tile, int& n, int& oh0, int& ow0, int& nt) {
    nt = tile / m_tiles;
    int m = tile % m_tiles;
    ow0 = (m % tiles_w) * 16;
    m /= tiles_w;
    oh0 = (m % tiles_h) * 8;
    n = m / tiles_h;
  };

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int it = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        int n, oh0, ow0, nt;
        decode_tile(tile, n, oh0, ow0, nt);
        for (int ki = 0; ki < k_iters; ++ki, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          const int tap = ki / k_chunks, c0 = (ki % k_chunks) * 64;
          const int kh = tap / p.taps_w, kw = tap % p.taps_w;
          const int ix = ow0 * p.stride + kw - p.pad, iy = oh0 * p.stride + kh - p.pad;
          uint8_t* st = smem + s * stage_bytes;
          mbar_arrive_expect_tx(&full[s], stage_bytes);
          tma_load_4d(st, &tmAhi, &full[s], c0, ix, iy, n);
          tma_load_3d(st + kATile, &tmBhi, &full[s], c0, nt * (p.n_pass == 2 ? 2 : 1) * p.BN, tap);
          if (p.n_pass == 3) {
            tma_load_4d(st + kATile + b_bytes, &tmAlo, &full[s], c0, ix, iy, n);
            tma_load_3d(st + 2 * kATile + b_bytes, &tmBlo, &full[s], c0, nt * p.BN, tap);
          }
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ UMMA issuer
    // whole warp, uniform control flow; descriptors advanced by one add per k-step; elected lane issues
    {
      const bool leader = elect_one();
      const uint32_t idesc = p.n_pass == 2 ? make_idesc(128, 2 * p.BN, false, false, kFmtF16, kFmtF16)
                                           : make_idesc_bf16(128, p.BN, false, false);
      int it = 0, local = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile, ++local) {
        const int acc = local & 1;
        mbar_wait(&tmem_empty[acc], ((local >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem + acc * 256;
        for (int ki = 0; ki < k_iters; ++ki, ++it) {
          const int s = it % p.stages;
          mbar_wait(&full[s], (it / p.stages) & 1);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + s * stage_bytes);
          const uint64_t dAh = desc_kmajor_sw128(a_hi), dBh = desc_kmajor_sw128(a_hi + kATile);
          const uint64_t dAl = desc_kmajor_sw128(a_hi + kATile + b_bytes), dBl = desc_kmajor_sw128(a_hi + 2 * kATile + b_bytes);
          if (leader) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t o = ks * 32;
              umma_ss(d, desc_advance(dAh, o), desc_advance(dBh, o), idesc, (ki | ks) != 0);
              if (p.n_pass == 3) {
                umma_ss(d, desc_advance(dAl, o), desc_advance(dBh, o), idesc, 1);
                umma_ss(d, desc_advance(dAh, o), desc_advance(dBl, o), idesc, 1);
              }
            }
            umma_commit(&empty[s]);
          }
        }
        if (leader) umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (0..3)
    const int r = threadIdx.x;  // TMEM lane = pixel inside the 8x16 tile
    const uint32_t lane_off = uint32_t(warp * 32) << 16;
    int local = 0;
    // GroupNorm statistics of the output: per-thread (= per-pixel-slot) running sums for every channel quad of the
    // current N tile, kept in registers across tiles and folded (warp shuffle -> shared -> double atomics) only when
    // the (image, N tile) pair changes — a handful of times per CTA thanks to the contiguous tile ranges.
    float st1[32], st2[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) st1[i] = st2[i] = 0.f;
    int st_n = -1, st_nt = -1;
    const int cpg = p.stats ? p.Cout / p.groups : 1;
    auto flush_stats = [&]() {
      if (st_n < 0) return;
#pragma unroll
      for (int qd = 0; qd < 32; ++qd) {
        if (qd * 4 >= p.BN) break;
        float s1 = st1[qd], s2 = st2[qd];
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) {
          s1 += __shfl_xor_sync(0xffffffffu, s1, sh);
          s2 += __shfl_xor_sync(0xffffffffu, s2, sh);
        }
        const int c = st_nt * p.BN + qd * 4;
        if (lane == 0 && c < p.Cout) {
          atomicAdd(&s_stats[2 * (c / cpg)], s1);
          atomicAdd(&s_stats[2 * (c / cpg) + 1], s2);
        }
        st1[qd] = st2[qd] = 0.f;
      }
      named_bar_sync(2, 128);
      if (r < 2 * p.groups) {
        const float t = s_stats[r];
        if (t != 0.f) atomicAdd(&p.stats[(size_t)st_n * p.groups * 2 + r], (double)t);
        s_stats[r] = 0.f;
      }
      named_bar_sync(2, 128);
    };
    for (int tile = tile_begin; tile < tile_end; ++tile, ++local) {
      int n, oh0, ow0, nt;
      decode_tile(tile, n, oh0, ow0, nt);
      if (p.stats && (n != st_n || nt != st_nt)) {
        flush_stats();
        st_n = n;
        st_nt = nt;
      }
      const int acc = local & 1;
      mbar_wait(&tmem_full[acc], (local >> 1) & 1);
      tc_fence_after();
      const size_t pix = ((size_t)n * p.Ho + oh0 + (r >> 4)) * p.Wo + ow0 + (r & 15);
      float* dst = p.out + pix * p.Cout;
      const float* res = p.residual ? p.residual + pix * p.Cout : nullptr;
      const int c_base = nt * p.BN;
#pragma unroll
      for (int ci = 0; ci < 16; ++ci) {
        const int c0 = ci * 16;
        if (c0 >= p.BN) break;
        uint32_t v[16];
        tmem_ld_x16(tmem + lane_off + acc * 256 + c0, v);
        if (p.n_pass == 2) {
          uint32_t w[16];
          tmem_ld_x16(tmem + lane_off + acc * 256 + p.BN + c0, w);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j)
            v[j] = __float_as_uint((__uint_as_float(v[j]) + __uint_as_float(w[j])) * p.w_scale_inv);
        } else {
          tmem_wait_ld();
        }
        const int c = c_base + c0;
        if (c + 16 <= p.Cout && (p.Cout & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(p.bias + c + j);
            float4 o = make_float4(__uint_as_float(v[j]) + b4.x, __uint_as_float(v[j + 1]) + b4.y,
                                   __uint_as_float(v[j + 2]) + b4.z, __uint_as_float(v[j + 3]) + b4.w);
            if (res) {
              const float4 r4 = *reinterpret_cast<const float4*>(res + c + j);
              o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
            }
            if (p.clip) {
              o.x = fminf(fmaxf(o.x, -1.f), 1.f); o.y = fminf(fmaxf(o.y, -1.f), 1.f);
              o.z = fminf(fmaxf(o.z, -1.f), 1.f); o.w = fminf(fmaxf(o.w, -1.f), 1.f);
            }
            *reinterpret_cast<float4*>(dst + c + j) = o;
            if (p.stats && ci < 8) {   // stats need BN <= 128 (checked on the host); a quad never straddles a group
              st1[ci * 4 + j / 4] += (o.x + o.y) + (o.z + o.w);
              st2[ci * 4 + j / 4] += (o.x * o.x + o.y * o.y) + (o.z * o.z + o.w * o.w);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c + j < p.Cout) {
              float o = __uint_as_float(v[j]) + p.bias[c + j];
              if (res) o += res[c + j];
              if (p.clip) o = fminf(fmaxf(o, -1.f), 1.f);
              dst[c + j] = o;
            }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
    }
    if (p.stats) flush_stats();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<512>(tmem);
}

}  // namespace lwm

using namespace lwm;

static int conv_launch(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, const float* bias,
                       const float* residual, float* out, int N, int Hin, int Win, int Cpad, int Ho, int Wo, int Cout,
                       int Cout_pad, int ksize, int stride, int pad, int n_pass, int clip, float w_scale_inv,
                       double* stats, int groups, void* stream) {
  if (!a_hi || !w_hi || !bias || !out) return lwm_fail(LWM_ERR_ARG, "vq_conv2d: null pointer");
  if (n_pass < 1 || n_pass > 3) return lwm_fail(LWM_ERR_ARG, "vq_conv2d: n_pass must be 1 (bf16), 2 (fp16x2) or 3 (bf16x3)");
  if (n_pass == 3 && (!a_lo || !w_lo)) return lwm_fail(LWM_ERR_ARG, "vq_conv2d: n_pass=3 needs the lo planes");
  if (Cpad % 64 || Cout_pad % 16 || Cout > Cout_pad) return lwm_fail(LWM_ERR_SHAPE, "vq_conv2d: Cpad % 64, Cout_pad % 16");
  if (Ho % 8 || Wo % 16) return lwm_fail(LWM_ERR_SHAPE, "vq_conv2d: output must tile by 8 x 16 pixels");
  if ((ksize != 1 && ksize != 3) || (stride != 1 && stride != 2)) return lwm_fail(LWM_ERR_SHAPE, "vq_conv2d: ksize 1|3, stride 1|2");
  if (stats && (groups <= 0 || groups > 64 || Cout % groups || (Cout / groups) % 4 || Cout % 16))
    return lwm_fail(LWM_ERR_SHAPE, "vq_conv2d: output statistics need Cout % 16 == 0 and (Cout/groups) % 4 == 0, groups <= 64");
  if (stats && n_pass != 2)
    return lwm_fail(LWM_ERR_ARG, "vq_conv2d: output statistics are an epilogue of the fp16x2 scheme (N tile <= 128)");
  if (!lwm_check_device()) return LWM_ERR_DEVICE;
  int BN = Cout_pad;
  if (n_pass == 2) {      // fp16x2 stacks hi|lo: the UMMA is 2*BN wide -> BN = largest multiple of 16 <= 128 dividing Cout_pad
    for (BN = 128; BN >= 16; BN -= 16)
      if (Cout_pad % BN == 0) break;
  } else if (BN > 256) {
    if (Cout_pad % 256 == 0) BN = 256;
    else if (Cout_pad % 192 == 0) BN = 192;
    else if (Cout_pad % 128 == 0) BN = 128;
    else return lwm_fail(LWM_ERR_SHAPE, "vq_conv2d: Cout_pad > 256 must be a multiple of 128");
  }
  const int taps = ksize * ksize;
  const CUtensorMapDataType dt = n_pass == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const int wrows = (n_pass == 2 ? 2 : 1);
  CUtensorMap tAh, tAl, tBh, tBl;
  {
    uint64_t dims[4] = {uint64_t(Cpad), uint64_t(Win), uint64_t(Hin), uint64_t(N)};
    uint64_t strides[3] = {uint64_t(Cpad) * 2, uint64_t(Win) * Cpad * 2, uint64_t(Hin) * Win * Cpad * 2};
    uint32_t box[4] = {64, uint32_t(16 * stride), uint32_t(8 * stride), 1};
    uint32_t es[4] = {1, uint32_t(stride), uint32_t(stride), 1};
    if (!encode_tmap(&tAh, dt, 4, a_hi, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, es))
      return lwm_fail(LWM_ERR_CUDA, "vq_conv2d: activation tensor map failed");
    if (n_pass == 3 &&
        !encode_tmap(&tAl, dt, 4, a_lo, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, es))
      return lwm_fail(LWM_ERR_CUDA, "vq_conv2d: activation (lo) tensor map failed");
  }
  {
    uint64_t dims[3] = {uint64_t(Cpad), uint64_t(Cout_pad) * wrows, uint64_t(taps)};
    uint64_t strides[2] = {uint64_t(Cpad) * 2, uint64_t(Cout_pad) * wrows * Cpad * 2};
    uint32_t box[3] = {64, uint32_t(BN * wrows), 1};
    if (!encode_tmap(&tBh, dt, 3, w_hi, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))
      return lwm_fail(LWM_ERR_CUDA, "vq_conv2d: weight tensor map failed");
    if (n_pass == 3 &&
        !encode_tmap(&tBl, dt, 3, w_lo, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))
      return lwm_fail(LWM_ERR_CUDA, "vq_conv2d: weight (lo) tensor map failed");
  }
  if (n_pass != 3) { tAl = tAh; tBl = tBh; }
  ConvParams p;
  p.N = N; p.Ho = Ho; p.Wo = Wo; p.Cout = Cout; p.Cpad = Cpad;
  p.taps_w = ksize; p.taps = taps; p.stride = stride; p.pad = pad;
  p.BN = BN; p.n_tiles = Cout_pad / BN; p.n_pass = n_pass;
  p.bias = bias; p.residual = residual; p.out = out; p.clip = clip;
  p.w_scale_inv = w_scale_inv; p.stats = stats; p.groups = groups;
  const int stage_bytes = (n_pass == 3 ? 2 : 1) * (kATile + wrows * BN * 128);
  int stages = (227 * 1024 - 2048) / stage_bytes;
  if (stages > 6) stages = 6;
  if (stages < 2) return lwm_fail(LWM_ERR_SHAPE, "vq_conv2d: tile does not fit in shared memory");
  p.stages = stages;
  const int smem_bytes = stages * stage_bytes + 1024;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  static bool attr_set_dev[64] = {};      // function attributes are per device
  if (!attr_set_dev[dev & 63]) {
    if (cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return lwm_fail(LWM_ERR_CUDA, "vq_conv2d: cannot raise dynamic shared memory limit");
    attr_set_dev[dev & 63] = true;
  }
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int total_tiles = N * (Ho / 8) * (Wo / 16) * p.n_tiles;
  const int grid = total_tiles < sms ? total_tiles : sms;
  conv_umma_kernel<<<grid, kConvThreads, smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(tAh, tAl, tBh, tBl, p);
  return lwm_check_launch("conv_umma_kernel");
}

// a_hi/a_lo: [N,Hin,Win,Cpad] bf16 planes; w_hi/w_lo: [taps][Cout_pad][Cpad] bf16; out/residual fp32 NHWC.
extern "C" int lwm_vq_conv2d(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo,
                             const float* bias, const float* residual, float* out, int N, int Hin, int Win,
                             int Cpad, int Ho, int Wo, int Cout, int Cout_pad, int ksize, int stride, int pad,
                             int n_pass, int clip, void* stream) {
  if (n_pass != 1 && n_pass != 3) return lwm_fail(LWM_ERR_ARG, "vq_conv2d: n_pass must be 1 (bf16) or 3 (bf16x3)");
  return conv_launch(a_hi, a_lo, w_hi, w_lo, bias, residual, out, N, Hin, Win, Cpad, Ho, Wo, Cout, Cout_pad, ksize,
                     stride, pad, n_pass, clip, 1.0f, nullptr, 0, stream);
}

// fp16x2 mode: a [N,Hin,Win,Cpad] fp16 plane (lwm_vq_prep_f16); w_stacked [taps][Cout_pad/BN][2*BN][Cpad] fp16 with
// BN = the largest multiple of 16 <= 128 that divides Cout_pad: rows [0,BN) = fp16(w / w_scale_inv), rows [BN,2BN) = fp16(w / w_scale_inv - hi).
// gn_stats_out (optional, zeroed by the caller): [N, groups, 2] double (sum, sumsq) of the OUTPUT tensor.
extern "C" int lwm_vq_conv2d_f16(const void* a, const void* w_stacked, const float* bias, const float* residual,
                                 float* out, double* gn_stats_out, int N, int Hin, int Win, int Cpad, int Ho, int Wo,
                                 int Cout, int Cout_pad, int ksize, int stride, int pad, float w_scale_inv, int groups,
                                 int clip, void* stream) {
  if (!(w_scale_inv > 0.f)) return lwm_fail(LWM_ERR_ARG, "vq_conv2d_f16: w_scale_inv must be positive");
  return conv_launch(a, nullptr, w_stacked, nullptr, bias, residual, out, N, Hin, Win, Cpad, Ho, Wo, Cout, Cout_pad,
                     ksize, stride, pad, 2, clip, w_scale_inv, gn_stats_out, groups, stream);
}
