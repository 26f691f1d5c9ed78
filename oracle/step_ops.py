"""CPU (torch float32/float64) stand-ins for the C-ABI step functions, with the same carry
semantics (include/lwm_b200.h: lwm_attn_fwd_step / lwm_attn_bwd_step). They let the ring
sequencing code (lwm_b200/ring_exec.py) run under the gloo backend in the CPU test-suite.
TEST INFRASTRUCTURE ONLY — never imported by the product path.

The arithmetic follows SURVEY.md Appendix A (online-softmax carry (numerator, denominator, max)
and the custom_vjp backward recurrences), evaluated densely per (q chunk, kv block) pair."""
import math

import torch

MASKED = -1.0e30
LOG2E = 1.4426950408889634


def _logits2(q, k, q_pos0, k_pos0, causal, bias, seg):
    """log2-domain logits [B,H,Sq,Sk] with masked entries set to MASKED (like the kernel)."""
    B, Sq, H, D = q.shape
    Sk = k.shape[1]
    t = torch.einsum("bqhd,bkhd->bhqk", q.double(), k.double()) * (LOG2E / math.sqrt(D))
    qp = q_pos0 + torch.arange(Sq)
    kp = k_pos0 + torch.arange(Sk)
    mask = torch.zeros(B, 1, Sq, Sk, dtype=torch.bool)
    if bias is not None:
        bt = bias[:, kp].double() * LOG2E
        t = t + bt[:, None, None, :]
        mask = mask | (bt < MASKED)[:, None, None, :]
    if seg is not None:
        mask = mask | (seg[:, qp][:, None, :, None] != seg[:, kp][:, None, None, :])
    if causal:
        mask = mask | (qp[:, None] < kp[None, :])[None, None]
    return torch.where(mask, torch.full_like(t, MASKED), t)


class CpuOps:
    @staticmethod
    def fwd_step(q, k, v, out, lse, acc_o, acc_m, acc_l, q_pos0, k_pos0, causal, bias, seg, first, last):
        t = _logits2(q, k, q_pos0, k_pos0, causal, bias, seg)
        m_loc = t.max(dim=-1).values                              # [B,H,Sq]
        p = torch.exp2(t - m_loc[..., None])
        l_loc = p.sum(-1)
        o_loc = torch.einsum("bhqk,bkhd->bqhd", p, v.double())
        if first:
            m_new, l_new, o_new = m_loc, l_loc, o_loc
        else:
            m_c, l_c, o_c = acc_m.double(), acc_l.double(), acc_o.double()
            m_new = torch.maximum(m_c, m_loc)
            wa, wb = torch.exp2(m_c - m_new), torch.exp2(m_loc - m_new)
            l_new = wa * l_c + wb * l_loc
            o_new = o_c * wa.transpose(1, 2)[..., None] + o_loc * wb.transpose(1, 2)[..., None]
        if last:
            out.copy_((o_new / l_new.transpose(1, 2)[..., None]).to(out.dtype))
            lse.copy_(((m_new + torch.log2(l_new)) / LOG2E).to(lse.dtype))
        else:
            acc_o.copy_(o_new.to(acc_o.dtype))
            acc_m.copy_(m_new.to(acc_m.dtype))
            acc_l.copy_(l_new.to(acc_l.dtype))

    @staticmethod
    def bwd_prep(out, dout, delta):
        delta.copy_((out.double() * dout.double()).sum(-1).transpose(1, 2).to(delta.dtype))

    @staticmethod
    def lse_for_bwd(lse):
        return lse            # the CPU stand-in keeps the natural-log lse

    @staticmethod
    def bwd_step(q, k, v, dout, lse, delta, dq_acc, dk_acc, dv_acc, q_pos0, k_pos0, causal, bias, seg):
        D = q.shape[-1]
        t = _logits2(q, k, q_pos0, k_pos0, causal, bias, seg)
        # rows whose every visited key was masked carry lse ~ MASKED/log2(e): their (arbitrary, padded)
        # forward value must not produce a gradient — and fp32 lse cannot resolve MASKED anyway
        dead = (lse.double() < -1.0e29)[..., None]
        p = torch.where(dead, torch.zeros_like(t), torch.exp2(t - lse.double()[..., None] * LOG2E))
        g = dout.double()
        dv_acc += torch.einsum("bhqk,bqhd->bkhd", p, g).to(dv_acc.dtype)
        dp = torch.einsum("bqhd,bkhd->bhqk", g, v.double())
        ds = p * (dp - delta.double()[..., None]) / math.sqrt(D)
        dq_acc += torch.einsum("bhqk,bkhd->bqhd", ds, k.double()).to(dq_acc.dtype)
        dk_acc += torch.einsum("bhqk,bqhd->bkhd", ds, q.double()).to(dk_acc.dtype)

    @staticmethod
    def cast(src, dst):
        dst.copy_(src.to(dst.dtype))

    @staticmethod
    def accumulate(acc, start, length, buf):
        acc[:, start:start + length] += buf
