"""On-GPU self-check used by __graft_entry__.smoke(): one small invocation of each hot path,
compared against the CPU oracle (the oracle is only ever the checker — see oracle/__init__.py)."""
import numpy as np
import torch


def _rel(x, ref):
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.linalg.norm(x - ref) / max(np.linalg.norm(ref), 1e-30))


def smoke(verbose=True):
    from oracle.attn_dense import attention_dense, attention_dense_grads
    from oracle import vqgan_ref as vr
    from . import ringattention as ra
    from .vqgan import Ops, PackedConv
    res = {}
    # ---- ring attention, forward + backward through the reference-signature op (ring size 1)
    g = torch.Generator().manual_seed(0)
    B, S, H, D = 1, 512, 2, 128
    q, k, v, do = [torch.randn(B, S, H, D, generator=g).to(torch.bfloat16).cuda() for _ in range(4)]
    q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
    out = ra.ringattention(q, k, v, None, None, axis_name="sp", float32_logits=True, cache_idx=None,
                           blockwise_kwargs=dict(causal_block_size=1, deterministic=True, attn_pdrop=0.0,
                                                 query_chunk_size=256, key_chunk_size=256))
    out.backward(do)
    torch.cuda.synchronize()
    n = lambda t: t.detach().float().cpu().numpy()  # noqa: E731
    ref = attention_dense(n(q), n(k), n(v), causal=True)
    rq, rk, rv = attention_dense_grads(n(q), n(k), n(v), n(do), causal=True)
    res["attn_fwd_rel_err"] = _rel(n(out), ref)
    res["attn_dq_rel_err"] = _rel(n(q.grad), rq)
    res["attn_dk_rel_err"] = _rel(n(k.grad), rk)
    res["attn_dv_rel_err"] = _rel(n(v.grad), rv)
    assert res["attn_fwd_rel_err"] < 3e-3 and max(res["attn_dq_rel_err"], res["attn_dk_rel_err"],
                                                  res["attn_dv_rel_err"]) < 5e-3, res
    # ---- VQGAN: GroupNorm+SiLU prep -> tcgen05 conv, and the nearest-code search (bit-exact)
    ops = Ops("bf16x3")
    x = torch.randn(1, 16, 16, 128, generator=g)
    gn = vr._gn_p(g, 128)
    cp = vr._conv_p(g, 3, 128, 128)
    y = ops.conv(ops.prep(x.cuda(), {"scale": gn["scale"].cuda(), "bias": gn["bias"].cuda()}),
                 PackedConv(cp, torch.device("cuda")))
    yref = vr.conv2d(vr.silu(vr.group_norm(x, gn)), cp)
    res["vqgan_conv_rel_err"] = _rel(n(y), yref.numpy())
    assert res["vqgan_conv_rel_err"] < 1e-3, res
    emb = torch.randn(8192, 64, generator=g)
    z = torch.randn(256, 64, generator=g)
    _, idx = ops.vq_argmin(z.cuda(), emb.cuda())
    _, ref_idx = vr.vector_quantize(z.numpy(), emb.numpy())
    torch.cuda.synchronize()
    res["vq_indices_bit_exact"] = bool(np.array_equal(idx.cpu().numpy().astype(np.int32), ref_idx))
    assert res["vq_indices_bit_exact"], "VQ indices differ from the oracle"
    if verbose:
        print("smoke:", {k2: (("%.2e" % v2) if isinstance(v2, float) else v2) for k2, v2 in res.items()})
    return res
