"""Shared helpers for the parity tests (seeded inputs, error metrics)."""
import numpy as np
import torch


def rel_fro(x, ref):
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.linalg.norm(x - ref) / max(np.linalg.norm(ref), 1e-30))


def make_qkv(B, Sq, Sk, H, D=128, seed=1234, device="cuda", n_extra=0):
    """N(0,1) generated in fp32 then rounded to bf16 (SURVEY.md §8d synthetic inputs)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    q = torch.randn(B, Sq, H, D, generator=g).to(torch.bfloat16)
    k = torch.randn(B, Sk, H, D, generator=g).to(torch.bfloat16)
    v = torch.randn(B, Sk, H, D, generator=g).to(torch.bfloat16)
    extra = [torch.randn(B, Sq, H, D, generator=g).to(torch.bfloat16) for _ in range(n_extra)]
    outs = [q, k, v] + extra
    return [t.to(device) for t in outs]


def to_np(t):
    return t.detach().float().cpu().numpy()
