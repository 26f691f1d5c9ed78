"""Step-for-step CPU restatement of the reference's ring x blockwise attention (forward and
custom_vjp backward), simulating the `sp` ranks of the ring in one process.

The arithmetic lives in the un-vendored pip dependency `ringattention`
(git+https://github.com/haoliuhl/ringattention.git, un-pinned: gpu_requirements.txt:8); it is bound
at lwm/llama.py:30 and called at lwm/llama.py:541-569. This file follows the published algorithm
as restated in SURVEY.md Appendix A (ringattention / blockwise_fwd / below_or_on_diag /
chunk_bias / bwd), with the call-site facts from lwm/llama.py:525-570: float32_logits=True,
causal_block_size=1, additive finfo.min masks combined with `minimum`, fp32 carries
(numerator, denominator, max_score), K/V rotated one hop per step (lax.ppermute i -> i+1).

numpy float32 throughout (the reference's fp32 logits / carries). TEST INFRASTRUCTURE ONLY.
PARITY UNPINNED against the reference binary — see oracle/__init__.py.
"""
import numpy as np

from .attn_dense import finfo_min


def below_or_on_diag(r, rb, c, cb, cbs):
    """Appendix A: does tile (q chunk r of size rb, k chunk c of size cb) touch the causal region?"""
    Q = max(cbs, rb)
    K = max(cbs, cb)
    r = r // (Q // rb)
    c = c // (K // cb)
    return ((r + 1) * Q - 1) > (c * K)


def chunk_bias(qi, kj, qc, kc, B, bias, seg, cbs, mask_value):
    """Appendix A chunk_bias: [B,1,qc,kc] float32, qi/kj are GLOBAL chunk indices."""
    q_off, k_off = qi * qc, kj * kc
    b = np.zeros((B, 1, 1, 1), dtype=np.float32)
    if bias is not None:
        b = np.asarray(bias, dtype=np.float32).reshape(B, 1, 1, -1)[:, :, :, k_off:k_off + kc]
    if seg is not None:
        seg = np.asarray(seg).reshape(B, -1)
        neq = seg[:, q_off:q_off + qc, None] != seg[:, None, k_off:k_off + kc]
        b = np.minimum(b, neq[:, None].astype(np.float32) * np.float32(mask_value))
    if cbs is not None:
        qp = (q_off + np.arange(qc)) // cbs
        kp = (k_off + np.arange(kc)) // cbs
        c = (qp[:, None] < kp[None, :]).astype(np.float32) * np.float32(mask_value)
        b = np.minimum(b, c[None, None])
    return np.broadcast_to(b, (B, 1, qc, kc)).astype(np.float32)


def _blockwise_fwd(q, k, v, carry, q0, k0, bias, seg, cbs, qc, kc, mask_value):
    num, den, mx = carry
    B, Sq, H, D = q.shape
    Sk = k.shape[1]
    scale = np.float32(1.0 / np.sqrt(D))
    for i in range(Sq // qc):
        qs = slice(i * qc, (i + 1) * qc)
        for j in range(Sk // kc):
            if cbs is not None and not below_or_on_diag(q0 + i, qc, k0 + j, kc, cbs):
                continue  # lax.cond skip of tiles entirely above the diagonal
            ks = slice(j * kc, (j + 1) * kc)
            s = np.einsum("bqhd,bkhd->bhqk", q[:, qs], k[:, ks]).astype(np.float32) * scale
            s = s + chunk_bias(q0 + i, k0 + j, qc, kc, B, bias, seg, cbs, mask_value)
            m_new = np.maximum(mx[:, :, qs], s.max(axis=-1))
            p = np.exp(s - m_new[..., None])
            c = np.exp(mx[:, :, qs] - m_new)
            num[:, qs] = num[:, qs] * c.transpose(0, 2, 1)[..., None] + np.einsum("bhqk,bkhd->bqhd", p, v[:, ks])
            den[:, :, qs] = den[:, :, qs] * c + p.sum(axis=-1)
            mx[:, :, qs] = m_new
    return num, den, mx


def ring_attention_fwd(q_shards, k_shards, v_shards, attn_bias=None, segment_ids=None, causal_block_size=1,
                       query_chunk_size=128, key_chunk_size=128, mask_dtype="bf16"):
    """q_shards/k_shards/v_shards: lists (one per sp rank) of [B,S_loc,H,D] arrays.
    Returns (outs, residuals) with outs[r] float32 [B,Sq,H,D]; residuals hold (den, mx) per rank."""
    P = len(q_shards)
    mask_value = finfo_min(mask_dtype)
    qc, kc = query_chunk_size, key_chunk_size
    qf = [np.asarray(x, dtype=np.float32) for x in q_shards]   # float32_logits=True: q,k upcast
    kf = [np.asarray(x, dtype=np.float32) for x in k_shards]
    vf = [np.asarray(x, dtype=np.float32) for x in v_shards]
    B, Sq, H, D = qf[0].shape
    Sk = kf[0].shape[1]
    carries = [(np.zeros((B, Sq, H, D), np.float32), np.zeros((B, H, Sq), np.float32),
                np.full((B, H, Sq), -np.inf, np.float32)) for _ in range(P)]
    held_k, held_v = list(kf), list(vf)
    with np.errstate(invalid="ignore"):
        for idx in range(P):
            for r in range(P):
                src = (r - idx) % P
                q0 = r * (Sq // qc)
                k0 = src * (Sk // kc)
                carries[r] = _blockwise_fwd(qf[r], held_k[r], held_v[r], carries[r], q0, k0, attn_bias,
                                            segment_ids, causal_block_size, qc, kc, mask_value)
            # ppermute i -> (i+1) mod P, every step including the last
            held_k = [held_k[(r - 1) % P] for r in range(P)]
            held_v = [held_v[(r - 1) % P] for r in range(P)]
    outs = [num / den.transpose(0, 2, 1)[..., None] for (num, den, mx) in carries]
    return outs, [(den, mx) for (num, den, mx) in carries]


def ring_attention_bwd(q_shards, k_shards, v_shards, outs, residuals, douts, attn_bias=None, segment_ids=None,
                       causal_block_size=1, query_chunk_size=128, key_chunk_size=128, mask_dtype="bf16"):
    """Appendix A `bwd`: returns (dq, dk, dv) lists of float32 arrays; dk/dv travel with k/v."""
    P = len(q_shards)
    mask_value = finfo_min(mask_dtype)
    qc, kc = query_chunk_size, key_chunk_size
    qf = [np.asarray(x, dtype=np.float32) for x in q_shards]
    kf = [np.asarray(x, dtype=np.float32) for x in k_shards]
    vf = [np.asarray(x, dtype=np.float32) for x in v_shards]
    gf = [np.asarray(x, dtype=np.float32) for x in douts]
    of = [np.asarray(x, dtype=np.float32) for x in outs]
    B, Sq, H, D = qf[0].shape
    Sk = kf[0].shape[1]
    scale = np.float32(1.0 / np.sqrt(D))
    dq = [np.zeros_like(x) for x in qf]
    held = [(kf[r], vf[r], np.zeros_like(kf[r]), np.zeros_like(vf[r])) for r in range(P)]
    for idx in range(P):
        for r in range(P):
            src = (r - idx) % P
            q0, k0 = r * (Sq // qc), src * (Sk // kc)
            k, v, dk, dv = held[r]
            den, mx = residuals[r]
            for i in range(Sq // qc):
                qs = slice(i * qc, (i + 1) * qc)
                delta = np.einsum("bqhd,bqhd->bhq", gf[r][:, qs], of[r][:, qs])
                for j in range(Sk // kc):
                    if causal_block_size is not None and not below_or_on_diag(q0 + i, qc, k0 + j, kc,
                                                                              causal_block_size):
                        continue
                    ks = slice(j * kc, (j + 1) * kc)
                    s = np.einsum("bqhd,bkhd->bhqk", qf[r][:, qs], k[:, ks]) * scale
                    s = s + chunk_bias(q0 + i, k0 + j, qc, kc, B, attn_bias, segment_ids, causal_block_size,
                                       mask_value)
                    p = np.exp(s - mx[:, :, qs, None]) / den[:, :, qs, None]
                    dv[:, ks] += np.einsum("bhqk,bqhd->bkhd", p, gf[r][:, qs])
                    dp = np.einsum("bqhd,bkhd->bhqk", gf[r][:, qs], v[:, ks])
                    dl = (dp - delta[..., None]) * p
                    dq[r][:, qs] += np.einsum("bhqk,bkhd->bqhd", dl, k[:, ks]) * scale
                    dk[:, ks] += np.einsum("bqhd,bhqk->bkhd", qf[r][:, qs], dl) * scale
        held = [held[(r - 1) % P] for r in range(P)]
    # after P hops every (k, v, dk, dv) tuple is back on its owner
    dk = [held[r][2] for r in range(P)]
    dv = [held[r][3] for r in range(P)]
    return dq, dk, dv


def torch_blockwise_fwd_bwd(q, k, v, dout, q_pos0=0, chunk=1024):
    """Same blockwise algorithm (online-softmax forward with (numerator, denominator, max) carry, then
    the recompute backward of Appendix A) for ONE device holding q rows [q_pos0, q_pos0+Sq) and all
    keys, written with torch CPU fp32 matmuls so that it can use every host core. Used by bench.py
    to time the reference algorithm on the host (cpu_baseline / --impl reference)."""
    import torch
    B, Sq, H, D = q.shape
    Sk = k.shape[1]
    scale = 1.0 / float(np.sqrt(D))
    qh, kh, vh, gh = [x.permute(0, 2, 1, 3).contiguous().float() for x in (q, k, v, dout)]  # [B,H,S,D]
    num = torch.zeros_like(qh)
    den = torch.zeros(B, H, Sq)
    mx = torch.full((B, H, Sq), -float("inf"))
    neg = float(np.finfo(np.float32).min)
    qc = min(chunk, Sq)
    for i in range(0, Sq, qc):
        qi = qh[:, :, i:i + qc]
        qpos = q_pos0 + i + torch.arange(qi.shape[2])
        for j in range(0, Sk, chunk):
            if j > q_pos0 + i + qc - 1:
                break                                   # tile entirely above the diagonal: skipped
            kj, vj = kh[:, :, j:j + chunk], vh[:, :, j:j + chunk]
            s = torch.matmul(qi, kj.transpose(-1, -2)) * scale
            kpos = j + torch.arange(kj.shape[2])
            if j + chunk - 1 > q_pos0 + i:
                s = s + (qpos[:, None] < kpos[None, :]).float() * neg
            m_new = torch.maximum(mx[:, :, i:i + qc], s.max(-1).values)
            p = torch.exp(s - m_new[..., None])
            c = torch.exp(mx[:, :, i:i + qc] - m_new)
            num[:, :, i:i + qc] = num[:, :, i:i + qc] * c[..., None] + torch.matmul(p, vj)
            den[:, :, i:i + qc] = den[:, :, i:i + qc] * c + p.sum(-1)
            mx[:, :, i:i + qc] = m_new
    out = num / den[..., None]
    dq = torch.zeros_like(qh)
    dk = torch.zeros_like(kh)
    dv = torch.zeros_like(vh)
    delta = (gh * out).sum(-1)
    for i in range(0, Sq, qc):
        qi, gi = qh[:, :, i:i + qc], gh[:, :, i:i + qc]
        qpos = q_pos0 + i + torch.arange(qi.shape[2])
        for j in range(0, Sk, chunk):
            if j > q_pos0 + i + qc - 1:
                break
            kj, vj = kh[:, :, j:j + chunk], vh[:, :, j:j + chunk]
            s = torch.matmul(qi, kj.transpose(-1, -2)) * scale
            kpos = j + torch.arange(kj.shape[2])
            if j + chunk - 1 > q_pos0 + i:
                s = s + (qpos[:, None] < kpos[None, :]).float() * neg
            p = torch.exp(s - mx[:, :, i:i + qc, None]) / den[:, :, i:i + qc, None]
            dv[:, :, j:j + chunk] += torch.matmul(p.transpose(-1, -2), gi)
            dp = torch.matmul(gi, vj.transpose(-1, -2))
            dl = (dp - delta[:, :, i:i + qc, None]) * p
            dq[:, :, i:i + qc] += torch.matmul(dl, kj) * scale
            dk[:, :, j:j + chunk] += torch.matmul(dl.transpose(-1, -2), qi) * scale
    back = lambda x: x.permute(0, 2, 1, 3)
    return back(out), back(dq), back(dk), back(dv)
