// Shared definitions for the ring-attention tile kernels (forward and backward).
#pragma once
#include "ptx.cuh"

namespace lwm {

constexpr int kHeadDim = 128;   // LWM-7B: hidden 4096 / 32 heads (lwm/llama.py:70-81)
constexpr int kTile = 128;      // rows per Q tile and keys per KV tile (one 128xN UMMA)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
// Logit (log2 domain) given to a masked position. The reference adds finfo(dtype).min, which
// absorbs the logit entirely in fp32, so every masked position carries the same value: a row
// whose visited keys are all masked averages them uniformly instead of producing NaN
// (SURVEY.md 8a "edge-case semantics"). A large finite constant reproduces exactly that.
constexpr float kMaskedLogit = -1.0e30f;

// Position-dependent inputs of one (q shard, kv block) step; everything is indexed by GLOBAL
// token position, as in the reference's _chunk_attention_bias.
struct MaskParams {
  int q_pos0;            // global position of local query row 0
  int k_pos0;            // global position of local key row 0
  int causal;            // 1 <=> causal_block_size == 1 ; 0 <=> causal_block_size is None
  const float* bias;     // [B, bias_stride] additive per-key bias at global key position, or null
  long long bias_stride;
  const int* seg;        // [B, seg_stride] segment ids at global position, or null
  long long seg_stride;
};

// Debug-only wait-time accounting (lwm_debug_set_prof): a role thread of CTA (0,0,0) accumulates the
// cycles it spends in each mbarrier wait; slot layout is documented next to each kernel.
struct WaitProf {
  unsigned long long* buf;   // null in production
  bool on;
  long long acc[12];
  LWM_DEVICE void init(unsigned long long* b) {
    buf = b;
    on = (b != nullptr) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) acc[i] = 0;
  }
  LWM_DEVICE void wait(uint64_t* bar, uint32_t parity, int slot) {
    if (on) {
      const long long t0 = clock64();
      mbar_wait(bar, parity);
      acc[slot] += clock64() - t0;
    } else {
      mbar_wait(bar, parity);
    }
  }
  LWM_DEVICE void flush(int base, int n, long long total) {
    if (on) {
      for (int i = 0; i < n; ++i) buf[base + i] = (unsigned long long)acc[i];
      buf[base + n] = (unsigned long long)total;
    }
  }
};

}  // namespace lwm
