"""CPU oracle of the rotary embedding on the attention prologue — TEST INFRASTRUCTURE ONLY (never imported by the
product). Restates lwm/llama.py:344-351 (`precompute_freqs_cis`) and :354-375 (`apply_rotary_emb`) in numpy, keeping
the reference's dtypes step by step: float32 frequencies, int64 x float32 outer product in float64 rounded to float32,
float32 sin/cos, complex64 table, complex64 multiply, final cast.

Parity status: PINNED — tests/golden/rope_reference.npz holds outputs of the reference's own two functions executed
here over the numpy-backed jax shim (tools/make_golden_next_rows_from_reference.py); tests/test_oracle_cpu.py checks
this restatement against them bit for bit."""
import numpy as np


def precompute_freqs_cis(dim, max_position_embedding, theta=10000.0, dtype=np.float32):
    """llama.py:344-351."""
    freqs = 1.0 / (theta ** (np.arange(0, dim, 2)[: (dim // 2)].astype(dtype) / dim))
    t = np.arange(max_position_embedding)
    freqs = np.outer(t, freqs).astype(dtype)
    sin, cos = np.sin(freqs), np.cos(freqs)
    return np.complex64(cos + 1j * sin)


def apply_rotary_emb(xq, xk, freqs_cis, dtype=np.float32):
    """llama.py:354-375. xq/xk [B,S,H,D] float arrays; freqs_cis [B,S,D/2] complex64 (table rows already gathered by
    position_ids, llama.py:515)."""
    def rot(x):
        x = np.asarray(x, dtype=np.float32).reshape(*x.shape[:-1], -1, 2)
        xc = (x[..., 0] + 1j * x[..., 1]).astype(np.complex64)
        f = freqs_cis.reshape(*freqs_cis.shape[:2], 1, *freqs_cis.shape[2:])     # add head dim
        y = xc * f
        return np.stack((np.real(y), np.imag(y)), axis=-1).reshape(*y.shape[:-1], -1).astype(dtype)
    return rot(xq), rot(xk)


def rope_reference(xq, xk, position_ids, theta, max_position, dtype=np.float32):
    """gather (llama.py:515) + apply."""
    table = precompute_freqs_cis(xq.shape[-1], max_position, theta)
    return apply_rotary_emb(xq, xk, np.take(table, position_ids, axis=0), dtype)
