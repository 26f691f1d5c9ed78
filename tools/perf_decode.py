"""Decode-op timing: one query row against a KV-cache shard (config 5: 1M tokens over 8 GPUs -> 131072 per rank)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lwm_b200.ringattention import ringattention_inference
for S in (16384, 131072):
    B, H, D = 1, 32, 128
    q = torch.randn(B, 1, H, D, device="cuda").to(torch.bfloat16)
    k = torch.randn(B, S, H, D, device="cuda").to(torch.bfloat16)
    v = torch.randn(B, S, H, D, device="cuda").to(torch.bfloat16)
    mask = torch.ones(B, 1, 1, S, dtype=torch.bool, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ringattention_inference(q, k, v, mask)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()                      # evict K/V from the 126 MB L2
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ringattention_inference(q, k, v, mask); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    gb = 2 * S * H * D * 2 / 1e9
    print("decode q_len=1, S_loc=%d: %.3f ms, %.0f GB/s of K/V streamed (%.2f of measured 6487 GB/s copy peak)" % (S, ms, gb / ms * 1e3, gb / ms * 1e3 / 6487.1))
