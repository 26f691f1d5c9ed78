"""Rotary position embedding of the attention prologue — host mirror of lwm/llama.py:344-375
(`precompute_freqs_cis`, `apply_rotary_emb`; used at llama.py:515-519) over the CUDA kernel `lwm_attn_rope`.

The reference precomputes a complex64 table [max_position, 64] on the host and gathers it by position_ids; the kernel
rebuilds the same float32 angles on the fly, so only the 64 inverse frequencies are kept on the device.
Differentiable: the VJP of a rotation is the rotation by the conjugate, done by the same kernel (conj=1).
No CPU path: tensors must live on a B200."""
import numpy as np
import torch

from . import _lib

_DT = {torch.float32: 0, torch.bfloat16: 1}


def precompute_inv_freq(dim, theta=10000.0, dtype=np.float32):
    """The `freqs` vector of precompute_freqs_cis (llama.py:345), same numpy expression and dtype."""
    return (1.0 / (theta ** (np.arange(0, dim, 2)[: (dim // 2)].astype(dtype) / dim))).astype(np.float32)


class RotaryTable:
    """Stands in for the reference's `freqs_cis` table: holds dim/theta (and max positions for range checks)."""

    def __init__(self, dim, max_position_embedding, theta=10000.0, device="cuda"):
        if dim != 128:
            raise _lib.LwmError("lwm_attn_rope is built for head_dim 128 (LWM-7B), got %d" % dim)
        self.dim, self.max_position, self.theta = dim, int(max_position_embedding), float(theta)
        self.inv_freq = torch.from_numpy(precompute_inv_freq(dim, theta)).to(device)


def precompute_freqs_cis(dim, max_position_embedding, theta=10000.0, dtype=np.float32, device="cuda"):
    """Same call shape as the reference (llama.py:344); returns a RotaryTable instead of a 512 MB complex array."""
    return RotaryTable(dim, max_position_embedding, theta, device)


def _launch(xq, xk, position_ids, table, out_dtype, conj):
    if not xq.is_cuda:
        raise _lib.LwmError("apply_rotary_emb: tensors must be CUDA tensors (no CPU path)")
    B, S, Hq, D = xq.shape
    Hk = xk.shape[2]
    if xk.shape[0] != B or xk.shape[1] != S or xk.shape[3] != D or xq.dtype != xk.dtype:
        raise _lib.LwmError("apply_rotary_emb: xq %s and xk %s disagree" % (tuple(xq.shape), tuple(xk.shape)))
    if xq.dtype not in _DT or out_dtype not in _DT:
        raise _lib.LwmError("apply_rotary_emb: dtypes must be float32 or bfloat16")
    if tuple(position_ids.shape) != (B, S):
        raise _lib.LwmError("apply_rotary_emb: position_ids must be [B,S]")
    xq, xk = xq.contiguous(), xk.contiguous()
    pos = position_ids.to(device=xq.device, dtype=torch.int32).contiguous()
    oq = torch.empty(xq.shape, dtype=out_dtype, device=xq.device)
    ok = torch.empty(xk.shape, dtype=out_dtype, device=xq.device)
    _lib.call("lwm_attn_rope", _lib.ptr(xq), _lib.ptr(xk), _DT[xq.dtype], _lib.ptr(oq), _lib.ptr(ok), _DT[out_dtype],
              _lib.ptr(pos), _lib.ptr(table.inv_freq), B, S, Hq, Hk, D, int(conj), _lib.stream_ptr())
    return oq, ok


class _Rope(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xq, xk, position_ids, table, out_dtype):
        ctx.table, ctx.in_dtype = table, xq.dtype
        ctx.save_for_backward(position_ids)
        return _launch(xq, xk, position_ids, table, out_dtype, conj=False)

    @staticmethod
    def backward(ctx, gq, gk):
        (position_ids,) = ctx.saved_tensors
        dq, dk = _launch(gq, gk, position_ids, ctx.table, ctx.in_dtype, conj=True)
        return dq, dk, None, None, None


def apply_rotary_emb(xq, xk, freqs_cis, dtype=torch.float32, *, position_ids):
    """apply_rotary_emb(xq, xk, freqs_cis, dtype) of llama.py:354-375 with the gather of llama.py:515 folded in:
    xq [B,S,Hq,128], xk [B,S,Hk,128] (head-split projections), freqs_cis = the RotaryTable from precompute_freqs_cis,
    position_ids [B,S] = the positions the reference gathers the table rows by. Returns (xq_out, xk_out) in `dtype`."""
    if not isinstance(freqs_cis, RotaryTable):
        raise _lib.LwmError("apply_rotary_emb: freqs_cis must come from lwm_b200.rope.precompute_freqs_cis")
    if int(position_ids.max()) >= freqs_cis.max_position or int(position_ids.min()) < 0:
        raise _lib.LwmError("apply_rotary_emb: position_ids outside [0, max_position_embedding)")
    return _Rope.apply(xq, xk, position_ids, freqs_cis, dtype)


def split_heads(x, num_heads, head_dim=128):
    """_split_heads (llama.py:376-377): [B,S,H*D] -> [B,S,H,D] — a view, never a copy."""
    return x.view(x.shape[0], x.shape[1], num_heads, head_dim)
