"""The public `ringattention` op on one GPU (ring size 1) in its DEFAULT precision mode against the float64 dense oracle,
and the small kernels the sharded executor is built from (shared-scale fp16 conversion, fused partial sum + cast,
delta from the fp16 dO copy, write-instead-of-accumulate dK/dV).

Tolerances: float32 inputs -> float32 results are the un-rounded accumulators: relative Frobenius error <= 1e-3
(north_star) for the forward and all three gradients; bfloat16 inputs -> bfloat16 results carry their own rounding
(8 significant bits, ~1.6e-3 rms): 3e-3."""
import numpy as np
import pytest
import torch

from helpers import make_qkv, rel_fro, to_np

pytestmark = pytest.mark.gpu
KW = dict(axis_name="sp", float32_logits=True, cache_idx=None,
          blockwise_kwargs=dict(causal_block_size=1, deterministic=True, attn_pdrop=0.0, query_chunk_size=256,
                                key_chunk_size=256))


@pytest.mark.parametrize("S,H", [(512, 2), (2048, 2)])
def test_default_mode_fp32_in_out_meets_1e3(S, H):
    from lwm_b200 import ringattention as ra
    from oracle.attn_dense import attention_dense, attention_dense_grads
    assert ra._DEFAULT_PRECISION == "fp16"
    g = torch.Generator().manual_seed(5)
    q, k, v, do = [torch.randn(1, S, H, 128, generator=g).cuda() for _ in range(4)]      # genuine fp32 values
    q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
    out = ra.ringattention(q, k, v, None, None, **KW)
    out.backward(do)
    torch.cuda.synchronize()
    assert out.dtype == torch.float32 and q.grad.dtype == torch.float32
    ref = attention_dense(to_np(q), to_np(k), to_np(v), causal=True)
    rq, rk, rv = attention_dense_grads(to_np(q), to_np(k), to_np(v), to_np(do), causal=True)
    for got, want in ((out, ref), (q.grad, rq), (k.grad, rk), (v.grad, rv)):
        assert rel_fro(to_np(got), want) < 1e-3


def test_default_mode_bf16_in_out():
    from lwm_b200 import ringattention as ra
    from oracle.attn_dense import attention_dense, attention_dense_grads
    q, k, v, do = make_qkv(1, 1024, 1024, 2, n_extra=1, seed=9)
    q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
    out = ra.ringattention(q, k, v, None, None, **KW)
    out.backward(do)
    torch.cuda.synchronize()
    assert out.dtype == torch.bfloat16 and k.grad.dtype == torch.bfloat16
    ref = attention_dense(to_np(q), to_np(k), to_np(v), causal=True)
    rq, rk, rv = attention_dense_grads(to_np(q), to_np(k), to_np(v), to_np(do), causal=True)
    for got, want in ((out, ref), (q.grad, rq), (k.grad, rk), (v.grad, rv)):
        assert rel_fro(to_np(got), want) < 3e-3
    # exact-in-fp32 callers see the un-rounded result: the same bf16 values passed as float32
    q2, k2, v2 = [t.detach().float().requires_grad_(True) for t in (q, k, v)]
    o2 = ra.ringattention(q2, k2, v2, None, None, **KW)
    o2.backward(do.float())
    torch.cuda.synchronize()
    for got, want in ((o2, ref), (q2.grad, rq), (k2.grad, rk), (v2.grad, rv)):
        assert rel_fro(to_np(got), want) < 1e-3


def test_mask_extent_is_checked():
    from lwm_b200 import ringattention as ra
    q, k, v = make_qkv(1, 256, 256, 1)
    with pytest.raises(ValueError):
        ra.ringattention(q, k, v, torch.zeros(1, 1, 1, 128, device="cuda"), None, **KW)
    with pytest.raises(ValueError):
        ra.ringattention(q, k, v, None, torch.zeros(1, 128, dtype=torch.int32, device="cuda"), **KW)


def test_shared_scale_conversion_and_helpers():
    from lwm_b200.ringattention import PeerOpsF16 as ops
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(2, 256, 2, 128, generator=g) * 3.0).cuda()
    xb = x.to(torch.bfloat16)
    table = torch.zeros(3, 4, dtype=torch.int32, device="cuda")     # three "ranks"
    ops.absmax(x, table[0, 1:2])
    ops.absmax(xb, table[1, 1:2])
    table[2, 1] = torch.tensor([1000.0]).view(torch.int32)[0]        # a peer with a much larger shard maximum
    scale = ops.make_scale(table, 1)
    torch.cuda.synchronize()
    assert float(table[0, 1:2].view(torch.float32)) == float(x.abs().max())
    assert float(table[1, 1:2].view(torch.float32)) == float(xb.float().abs().max())
    assert float(scale) == 2.0 ** (9 - 12)                          # 1000 = 1.95 * 2^9
    y = torch.empty(x.shape, dtype=torch.float16, device="cuda")
    ops.stage(xb, y, scale)
    torch.cuda.synchronize()
    assert torch.equal(y.float() * float(scale), xb.float())        # bf16 -> scaled fp16 is exact
    ops.stage(x, y, scale)
    torch.cuda.synchronize()
    assert torch.equal(y, (x / float(scale)).to(torch.float16))     # fp32 -> one rounding to fp16
    # fused sum + cast
    parts = [torch.randn(1024, 128, generator=g).cuda() for _ in range(5)]
    want = parts[0] + parts[1] + parts[2] + parts[3] + parts[4]
    d32 = torch.empty_like(want)
    d16 = torch.empty(want.shape, dtype=torch.bfloat16, device="cuda")
    ops.reduce_cast(parts, d32)
    ops.reduce_cast(parts, d16)
    torch.cuda.synchronize()
    assert torch.equal(d32, want) and torch.equal(d16, want.to(torch.bfloat16))
    # delta from the scaled fp16 copy of dO
    out = torch.randn(1, 256, 2, 128, generator=g).cuda()
    do = torch.randn(1, 256, 2, 128, generator=g).to(torch.bfloat16).cuda()
    tb = torch.zeros(1, 4, dtype=torch.int32, device="cuda")
    ops.absmax(do, tb[0, 3:4])
    sdo = ops.make_scale(tb, 3)
    d16v = torch.empty(do.shape, dtype=torch.float16, device="cuda")
    ops.stage(do, d16v, sdo)
    for o in (out, out.to(torch.bfloat16)):
        delta = torch.empty(1, 2, 256, dtype=torch.float32, device="cuda")
        ops.bwd_prep(o, d16v, sdo, delta)
        torch.cuda.synchronize()
        ref = (o.double() * do.double()).sum(-1).transpose(1, 2)
        assert float((delta.double() - ref).abs().max()) < 1e-4


def test_bwd_init_writes_instead_of_accumulating():
    from lwm_b200 import ringattention as ra
    q, k, v, do = make_qkv(1, 512, 512, 2, n_extra=1, seed=3)
    out = torch.empty_like(q)
    lse = torch.empty(1, 2, 512, dtype=torch.float32, device="cuda")
    ra.fwd_step(q, k, v, out, lse, None, None, None, 0, 0, True, None, None, True, True)
    delta = torch.empty_like(lse)
    ra.bwd_prep(out, do, delta)
    nl = ra.lse_for_bwd(lse)
    acc = [torch.zeros(1, 512, 2, 128, dtype=torch.float32, device="cuda") for _ in range(3)]
    ra.bwd_step(q, k, v, do, nl, delta, *acc, 0, 0, True, None, None)
    junk = [torch.full((1, 512, 2, 128), 7.0, dtype=torch.float32, device="cuda") for _ in range(2)]
    dq2 = torch.zeros_like(acc[0])
    ra.bwd_step(q, k, v, do, nl, delta, dq2, *junk, 0, 0, True, None, None, init=True)
    torch.cuda.synchronize()
    assert torch.equal(junk[0], acc[1]) and torch.equal(junk[1], acc[2])
    # key tiles no query can see (q shard entirely before the keys) are zero-filled by an initialising launch
    junk = [torch.full((1, 512, 2, 128), 7.0, dtype=torch.float32, device="cuda") for _ in range(2)]
    ra.bwd_step(q[:, :128].contiguous(), k, v, do[:, :128].contiguous(), nl[:, :, :128].contiguous(),
                delta[:, :, :128].contiguous(), torch.zeros(1, 128, 2, 128, device="cuda"), *junk, 0, 0, True, None, None,
                init=True)
    torch.cuda.synchronize()
    assert float(junk[0][:, 128:].abs().max()) == 0.0 and float(junk[1][:, 128:].abs().max()) == 0.0
    assert float(junk[0][:, :128].abs().max()) > 0.0
