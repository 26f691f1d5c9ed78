"""On-GPU self-check used by __graft_entry__.smoke(): one small invocation of each hot path,
compared against the CPU oracle (the oracle is only ever the checker — see oracle/__init__.py)."""
import numpy as np
import torch


def _rel(x, ref):
    return float(np.linalg.norm(x - ref) / max(np.linalg.norm(ref), 1e-30))


def smoke(verbose=True):
    from oracle.attn_dense import attention_dense
    from . import ringattention as ra
    g = torch.Generator().manual_seed(0)
    B, S, H, D = 1, 512, 2, 128
    q, k, v = [torch.randn(B, S, H, D, generator=g).to(torch.bfloat16).cuda() for _ in range(3)]
    out = ra.ringattention(q, k, v, None, None, axis_name="sp", float32_logits=True, cache_idx=None,
                           blockwise_kwargs=dict(causal_block_size=1, deterministic=True, attn_pdrop=0.0,
                                                 query_chunk_size=256, key_chunk_size=256))
    torch.cuda.synchronize()
    ref = attention_dense(q.float().cpu().numpy(), k.float().cpu().numpy(), v.float().cpu().numpy(), causal=True)
    err = _rel(out.float().cpu().numpy(), ref)
    if verbose:
        print("smoke: ringattention fwd rel-Frobenius error vs oracle = %.3e" % err)
    assert err < 2e-3, err
    return {"attn_fwd_rel_err": err}
