/* lwm_b200 — C ABI of the B200 (sm_100a) hot-path library: liblwm_b200.so
 *
 * This is the drop-in boundary for the two LWM hot paths:
 *   - the blockwise `ringattention(q, k, v, attn_bias, segment_ids, ...)` call bound at
 *     /root/reference lwm/llama.py:539-569 (forward) and its custom_vjp backward;
 *   - the VQGAN tokenizer ops of lwm/vqgan.py:105-351 (conv / GroupNorm / SiLU / resample /
 *     nearest-codebook lookup).
 * The reference has no native boundary of its own (it is pure Python/JAX); these entry points are
 * what a Python (ctypes / XLA-FFI custom call) binding on the reference side would bind — see
 * INTEGRATION.md for the stub.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *   - all tensor memory is caller-owned; calls are asynchronous on `stream` (a cudaStream_t);
 *   - return value: 0 on success, LWM_ERR_* otherwise; lwm_last_error() gives the message
 *     (thread-local). There is NO CPU fallback: on a non-sm_100 device every call fails with
 *     LWM_ERR_DEVICE;
 *   - one host thread per GPU/process (torchrun model); contexts are not thread-safe.
 */
#ifndef LWM_B200_H_
#define LWM_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define LWM_B200_ABI_VERSION 1

#define LWM_OK 0
#define LWM_ERR_DEVICE 1
#define LWM_ERR_SHAPE 2
#define LWM_ERR_ARG 3
#define LWM_ERR_CUDA 4

int lwm_abi_version(void);
const char* lwm_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * Ring attention, forward: ONE ring step (one held K/V block against the local query shard).
 * Replaces the body of the reference's per-step `_blockwise_attention_fwd` (un-vendored
 * `ringattention` package; call site lwm/llama.py:541-569; algorithm SURVEY.md Appendix A).
 *
 *   q            [B, Sq, H, D] bf16      local query shard (contiguous)
 *   k, v         [B, Sk, H, D] bf16      the K/V block currently held by this rank
 *   out          [B, Sq, H, D] bf16      final output, written when `last`
 *   lse          [B, H, Sq]    fp32      log-sum-exp (natural log) of the scaled+biased logits,
 *                                        written when `last` (residual for the backward)
 *   acc_o/m/l    fp32 carries [B,Sq,H,D] / [B,H,Sq] / [B,H,Sq] (numerator, running max in the
 *                log2 domain, denominator) — the reference's (numerator, max_score, denominator)
 *                scan carry. Read unless `first`, written unless `last`. May be NULL when
 *                first && last (single-step, ring_size 1).
 *   q_pos0/k_pos0 global token position of local row 0 of q / of the held k block: causal,
 *                bias and segment masks are evaluated on GLOBAL positions as in the reference's
 *                _chunk_attention_bias.
 *   causal       1 <=> blockwise_kwargs.causal_block_size == 1 ; 0 <=> None
 *   bias         [B, bias_stride] fp32 additive per-key bias indexed by global key position
 *                (the [B,1,1,S_global] attn_bias of lwm/llama.py:533-537 squeezed), or NULL
 *   segment_ids  [B, seg_stride] int32 indexed by global position, or NULL
 *   softmax_scale 1/sqrt(D)
 * Constraints: D == 128; Sq, Sk multiples of 128.
 */
int lwm_attn_fwd_step(const void* q, const void* k, const void* v, void* out, float* lse, float* acc_o,
                      float* acc_m, float* acc_l, int B, int H, int Sq, int Sk, int D, long long q_pos0,
                      long long k_pos0, int causal, const float* bias, long long bias_stride,
                      const int* segment_ids, long long seg_stride, float softmax_scale, int first, int last,
                      void* stream);

/* Ring attention, backward.
 * lwm_attn_bwd_prep: delta[b,h,s] = sum_d dout[b,s,h,d] * out[b,s,h,d]  (fp32), once per backward.
 * lwm_attn_bwd_step: one ring step of the reference's custom_vjp bwd (SURVEY.md Appendix A `bwd`):
 *   recomputes P from (q, k, lse), accumulates
 *     dq_acc [B,Sq,H,D] fp32 += dS K / sqrt(D)        (atomic fp32 tile reductions; zero it first)
 *     dk_acc [B,Sk,H,D] fp32 += dS^T Q / sqrt(D)      (read-modify-write by the owning CTA)
 *     dv_acc [B,Sk,H,D] fp32 += P^T dO
 *   dk_acc/dv_acc travel with the K/V block around the ring exactly like the reference's dk, dv.
 */
int lwm_attn_bwd_prep(const void* out, const void* dout, float* delta, int B, int H, int Sq, int D, void* stream);
int lwm_attn_bwd_step(const void* q, const void* k, const void* v, const void* dout, const float* lse,
                      const float* delta, float* dq_acc, float* dk_acc, float* dv_acc, int B, int H, int Sq,
                      int Sk, int D, long long q_pos0, long long k_pos0, int causal, const float* bias,
                      long long bias_stride, const int* segment_ids, long long seg_stride, float softmax_scale,
                      void* stream);

/* Element-wise helpers used by the ring host loop. */
int lwm_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream);
/* dst[i] += src[i] (fp32, n % 4 == 0): folds a dK/dV partial received from a peer into the owner's accumulator. */
int lwm_add_f32(float* dst, const float* src, long long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LWM_B200_H_ */
