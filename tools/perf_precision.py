"""Whole-op cost of the two precision modes at the headline shape: ringattention fwd+bwd (autograd), 1 GPU,
q/k/v [1,S,32,128] bf16, causal. precision='bf16' (default) vs 'fp16' (the <= 1e-3 mode: operand conversion passes,
fp32 output residual). Usage: python tools/perf_precision.py [S]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lwm_b200.ringattention import ringattention  # noqa: E402


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    B, H, D = 1, 32, 128
    g = torch.Generator(device="cuda").manual_seed(1234)
    q, k, v, do = [torch.randn(B, S, H, D, device="cuda", generator=g).to(torch.bfloat16) for _ in range(4)]
    flops = 3.5 * 4.0 * B * H * D * S * (S + 1) / 2
    for prec in ("bf16", "fp16"):
        def step():
            qq, kk, vv = [t.detach().requires_grad_(True) for t in (q, k, v)]
            out = ringattention(qq, kk, vv, None, None, axis_name="sp", float32_logits=True, cache_idx=None,
                                blockwise_kwargs=dict(causal_block_size=1, deterministic=True, dropout_rng=None,
                                                      attn_pdrop=0.0, query_chunk_size=1024, key_chunk_size=1024,
                                                      dtype=torch.bfloat16, policy=None, precision=None,
                                                      prevent_cse=True), precision=prec)
            out.backward(do)
        for _ in range(2):
            step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 3
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print("precision=%s S=%d: %.1f ms/step fwd+bwd, %.0f TFLOP/s" % (prec, S, ms, flops / ms / 1e9), flush=True)


if __name__ == "__main__":
    main()
