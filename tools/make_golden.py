"""Regenerates tests/golden/*.npz from the CPU oracle (float64 dense attention, numpy VQ) on tiny
seeded inputs. The reference itself cannot be imported here (no jax/flax offline) and ships no
golden vectors, so these fixtures pin the ORACLE against regressions and let the GPU box check the
kernels against committed numbers without recomputing the oracle. Run: python tools/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def bf16_round(x):
    """round-to-nearest-even to bfloat16, returned as float32 (numpy only)."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return u.astype(np.uint32).view(np.float32)


def main():
    from oracle.attn_dense import attention_dense, attention_dense_grads, finfo_min
    from oracle.vqgan_ref import vector_quantize
    out = os.path.join(ROOT, "tests", "golden")
    rng = np.random.default_rng(20260922)
    B, S, H, D = 1, 256, 2, 128
    q, k, v, do = [bf16_round(rng.standard_normal((B, S, H, D))) for _ in range(4)]
    bias = np.zeros((B, S), np.float32)
    bias[0, :19] = finfo_min("bf16")
    seg = np.zeros((B, S), np.int32)
    seg[0, 150:] = 1
    o, lse = attention_dense(q, k, v, causal=True, return_lse=True)
    dq, dk, dv = attention_dense_grads(q, k, v, do, causal=True)
    om, _ = attention_dense(q, k, v, causal=True, attn_bias=bias, segment_ids=seg, return_lse=True)
    np.savez_compressed(os.path.join(out, "attn_s256.npz"), q=q, k=k, v=v, do=do, bias=bias, seg=seg,
                        out=o.astype(np.float32), lse=lse.astype(np.float32), dq=dq.astype(np.float32),
                        dk=dk.astype(np.float32), dv=dv.astype(np.float32), out_masked=om.astype(np.float32))
    emb = rng.standard_normal((8192, 64)).astype(np.float32)
    z = rng.standard_normal((256, 64)).astype(np.float32)
    zq, idx = vector_quantize(z, emb)
    np.savez_compressed(os.path.join(out, "vq_256x8192.npz"), z=z, emb_seed=np.int64(20260922), idx=idx,
                        zq_checksum=np.float64(zq.astype(np.float64).sum()), emb_checksum=np.float64(emb.astype(np.float64).sum()))
    print("wrote", os.listdir(out))


if __name__ == "__main__":
    main()
