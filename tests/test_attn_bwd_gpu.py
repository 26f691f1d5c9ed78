"""GPU parity of the backward ring-attention tile kernel against the float64 closed-form
gradients of the dense oracle (oracle/attn_dense.py), same bf16-rounded inputs.

Tolerance: relative Frobenius error of the fp32 gradient accumulators <= 3e-3 for dq/dk/dv on
white-noise inputs. P and dS are rounded to bf16 (8 significant bits, rms relative rounding error
1.6e-3) before the tensor-core products, as in every bf16 flash-attention backward; on N(0,1)
inputs the result is itself a random-walk sum, so that rounding noise does not average out
relative to the signal (measured: dv 1.3e-3, dq/dk 2.2e-3). See DESIGN.md "Numerics"."""
import numpy as np
import pytest
import torch

from helpers import make_qkv, rel_fro, to_np

pytestmark = pytest.mark.gpu
TOL_GRAD = 3e-3


def _run_bwd(q, k, v, do, causal=True, q_pos0=0, k_pos0=0, bias=None, seg=None):
    from lwm_b200 import ringattention as ra
    B, Sq, H, D = q.shape
    Sk = k.shape[1]
    out = torch.empty_like(q)
    lse = torch.empty(B, H, Sq, dtype=torch.float32, device="cuda")
    ra.fwd_step(q, k, v, out, lse, None, None, None, q_pos0, k_pos0, causal, bias, seg, True, True)
    delta = torch.empty_like(lse)
    ra.bwd_prep(out, do, delta)
    dq = torch.zeros(B, Sq, H, D, dtype=torch.float32, device="cuda")
    dk = torch.zeros(B, Sk, H, D, dtype=torch.float32, device="cuda")
    dv = torch.zeros(B, Sk, H, D, dtype=torch.float32, device="cuda")
    ra.bwd_step(q, k, v, do, ra.lse_for_bwd(lse), delta, dq, dk, dv, q_pos0, k_pos0, causal, bias, seg)
    torch.cuda.synchronize()
    return out, lse, delta, dq, dk, dv


@pytest.mark.parametrize("B,S,H", [(1, 128, 1), (1, 256, 2), (2, 384, 2), (1, 1024, 2)])
@pytest.mark.parametrize("causal", [True, False])
def test_bwd_single_step(B, S, H, causal):
    from oracle.attn_dense import attention_dense_grads
    q, k, v, do = make_qkv(B, S, S, H, n_extra=1)
    out, lse, delta, dq, dk, dv = _run_bwd(q, k, v, do, causal)
    rq, rk, rv = attention_dense_grads(to_np(q), to_np(k), to_np(v), to_np(do), causal=causal)
    for name, got, ref in (("dq", dq, rq), ("dk", dk, rk), ("dv", dv, rv)):
        g = to_np(got)
        assert np.isfinite(g).all(), name
        assert rel_fro(g, ref) < TOL_GRAD, (name, rel_fro(g, ref))


def test_bwd_prep_delta():
    B, S, H = 2, 256, 3
    q, k, v, do = make_qkv(B, S, S, H, n_extra=1, seed=9)
    from lwm_b200 import ringattention as ra
    delta = torch.empty(B, H, S, dtype=torch.float32, device="cuda")
    ra.bwd_prep(q, do, delta)   # any two [B,S,H,D] bf16 tensors
    torch.cuda.synchronize()
    ref = (to_np(q).astype(np.float64) * to_np(do)).sum(-1).transpose(0, 2, 1)
    assert np.abs(to_np(delta) - ref).max() < 1e-3


def test_bwd_offsets_qlen_ne_kvlen():
    from oracle.attn_dense import attention_dense_grads
    B, Sq, Sk, H = 1, 256, 640, 2
    q, k, v, do = make_qkv(B, Sq, Sk, H, n_extra=1, seed=13)
    out, lse, delta, dq, dk, dv = _run_bwd(q, k, v, do, True, q_pos0=256, k_pos0=0)
    rq, rk, rv = attention_dense_grads(to_np(q), to_np(k), to_np(v), to_np(do), causal=True, q_pos0=256, k_pos0=0)
    assert rel_fro(to_np(dq), rq) < TOL_GRAD
    assert rel_fro(to_np(dk), rk) < TOL_GRAD
    assert rel_fro(to_np(dv), rv) < TOL_GRAD
    # keys beyond every query position receive exactly zero gradient
    assert np.abs(to_np(dk)[:, 512:]).max() == 0.0 and np.abs(to_np(dv)[:, 512:]).max() == 0.0


def test_bwd_bias_and_segments():
    from oracle.attn_dense import attention_dense_grads, finfo_min
    B, S, H = 1, 512, 2
    q, k, v, do = make_qkv(B, S, S, H, n_extra=1, seed=17)
    bias = torch.zeros(B, S, dtype=torch.float32)
    bias[0, :70] = finfo_min("bf16")
    seg = torch.zeros(B, S, dtype=torch.int32)
    seg[0, 300:] = 1
    # zero the cotangent of padded query rows: their forward value is arbitrary in the reference
    do = do.clone()
    do[:, :70] = 0
    out, lse, delta, dq, dk, dv = _run_bwd(q, k, v, do, True, bias=bias.cuda(), seg=seg.cuda())
    rq, rk, rv = attention_dense_grads(to_np(q), to_np(k), to_np(v), to_np(do), causal=True,
                                       attn_bias=bias.numpy(), segment_ids=seg.numpy())
    assert np.isfinite(to_np(dq)).all()
    assert rel_fro(to_np(dq)[:, 70:], rq[:, 70:]) < TOL_GRAD
    assert rel_fro(to_np(dk), rk) < TOL_GRAD
    assert rel_fro(to_np(dv), rv) < TOL_GRAD


def test_autograd_entry_point_matches_oracle():
    """the public op (reference signature) end to end, ring size 1."""
    from lwm_b200.ringattention import ringattention
    from oracle.attn_dense import attention_dense, attention_dense_grads
    B, S, H = 1, 512, 2
    q, k, v, do = make_qkv(B, S, S, H, n_extra=1, seed=23)
    q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
    out = ringattention(q, k, v, None, None, axis_name="sp", float32_logits=True, cache_idx=None,
                        blockwise_kwargs=dict(causal_block_size=1, deterministic=True, dropout_rng=None,
                                              attn_pdrop=0.0, query_chunk_size=128, key_chunk_size=256,
                                              dtype=torch.bfloat16, policy=None, precision=None, prevent_cse=True))
    out.backward(do)
    torch.cuda.synchronize()
    ref = attention_dense(to_np(q), to_np(k), to_np(v), causal=True)
    rq, rk, rv = attention_dense_grads(to_np(q), to_np(k), to_np(v), to_np(do), causal=True)
    assert rel_fro(to_np(out), ref) < 3e-3
    # bf16 gradients: bf16 rounding of the result itself (~1.6e-3 rms) on top of TOL_GRAD
    assert rel_fro(to_np(q.grad), rq) < 5e-3
    assert rel_fro(to_np(k.grad), rk) < 5e-3
    assert rel_fro(to_np(v.grad), rv) < 5e-3


def test_bwd_linearity_large():
    """size-independent property at a larger size: the backward is linear in dout."""
    B, S, H = 1, 4096, 4
    q, k, v, do = make_qkv(B, S, S, H, n_extra=1, seed=29)
    _, _, _, dq1, dk1, dv1 = _run_bwd(q, k, v, do)
    _, _, _, dq2, dk2, dv2 = _run_bwd(q, k, v, (2 * do.float()).to(torch.bfloat16))
    for a, b2 in ((dq1, dq2), (dk1, dk2), (dv1, dv2)):
        assert rel_fro(to_np(b2), 2 * to_np(a)) < 1e-3


def test_bwd_kv_split_additivity_large():
    """Size-independent property at a larger size: the backward over a K/V block equals the sum of the backwards
    over its two halves (dq accumulates across ring steps; dk/dv rows are disjoint)."""
    from lwm_b200 import ringattention as ra
    B, S, H = 1, 8192, 4
    q, k, v, do = make_qkv(B, S, S, H, n_extra=1, seed=53)
    out, lse, delta, dq, dk, dv = _run_bwd(q, k, v, do)
    nl = ra.lse_for_bwd(lse)
    dq2 = torch.zeros_like(dq)
    dk2, dv2 = torch.zeros_like(dk), torch.zeros_like(dv)
    h = S // 2
    for lo in (h, 0):     # ring order: later block first
        kk, vv = k[:, lo:lo + h].contiguous(), v[:, lo:lo + h].contiguous()
        dkh = torch.zeros(B, h, H, 128, dtype=torch.float32, device="cuda")
        dvh = torch.zeros_like(dkh)
        ra.bwd_step(q, kk, vv, do, nl, delta, dq2, dkh, dvh, 0, lo, True, None, None)
        dk2[:, lo:lo + h] = dkh
        dv2[:, lo:lo + h] = dvh
    torch.cuda.synchronize()
    assert rel_fro(to_np(dq2), to_np(dq)) < 1e-4        # same products, different atomic-add order
    assert rel_fro(to_np(dk2), to_np(dk)) < 1e-5
    assert rel_fro(to_np(dv2), to_np(dv)) < 1e-5


def test_fp32_inputs_take_the_documented_cast_path_in_bf16_mode():
    """precision='bf16' (the legacy operand mode): fp32 q/k/v get one rounding to bf16 on entry, fp32 output and
    gradients; must equal the bf16-input call on the pre-rounded tensors. (The default fp16 mode takes fp32 inputs
    natively: tests/test_attn_public_op_gpu.py.)"""
    from lwm_b200.ringattention import ringattention
    q, k, v, do = make_qkv(1, 256, 256, 2, n_extra=1)
    q32, k32, v32 = [(t.float() + 1e-4).requires_grad_(True) for t in (q, k, v)]   # not exactly representable in bf16
    qb, kb, vb = [t.detach().to(torch.bfloat16).requires_grad_(True) for t in (q32, k32, v32)]
    kw = dict(axis_name="sp", blockwise_kwargs=dict(causal_block_size=1), precision="bf16")
    o32 = ringattention(q32, k32, v32, None, None, **kw)
    ob = ringattention(qb, kb, vb, None, None, **kw)
    o32.backward(do.float())
    ob.backward(do)
    torch.cuda.synchronize()
    assert o32.dtype == torch.float32 and q32.grad.dtype == torch.float32
    assert torch.equal(o32, ob.float())
    for a, b in ((q32.grad, qb.grad), (k32.grad, kb.grad), (v32.grad, vb.grad)):
        assert torch.equal(a, b.float())
