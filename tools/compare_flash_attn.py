"""Independent cross-check and library baseline on the GPU box (NOT YET RUN — written after the round's GPU budget was
spent): flash_attn 2.8 (mma.sync-era kernels recompiled for this GPU; library code) against the lwm_b200
attention op on the same bf16 inputs, causal, [1,S,32,128].
  * numerics at S=4096: relative Frobenius distance of out / dq / dk / dv between the two implementations
    (both are bf16-P flash attention, so ~2e-3 is expected; an outlier flags a layout / scaling / mask bug);
  * speed at S=16384 and S=131072: fwd+bwd ms and TFLOP/s (causal algorithmic FLOPs) for both.
Usage: python tools/compare_flash_attn.py [S ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, warm=2, iters=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    from lwm_b200.ringattention import ringattention
    try:
        from flash_attn import flash_attn_func
    except Exception as e:                       # noqa: BLE001
        print("flash_attn not importable on this box: %r" % (e,))
        return
    sizes = [int(x) for x in sys.argv[1:]] or [4096, 16384, 131072]
    H, D = 32, 128
    kw = dict(axis_name="sp", blockwise_kwargs=dict(causal_block_size=1))
    for S in sizes:
        g = torch.Generator(device="cuda").manual_seed(1)
        q, k, v, do = [torch.randn(1, S, H, D, device="cuda", generator=g).to(torch.bfloat16) for _ in range(4)]

        def ours():
            a, b, c = [t.detach().requires_grad_(True) for t in (q, k, v)]
            o = ringattention(a, b, c, None, None, **kw)
            o.backward(do)
            return o, a.grad, b.grad, c.grad

        def flash():
            a, b, c = [t.detach().requires_grad_(True) for t in (q, k, v)]
            o = flash_attn_func(a, b, c, causal=True)
            o.backward(do)
            return o, a.grad, b.grad, c.grad
        try:
            ro, rf = ours(), flash()
        except Exception as e:                   # noqa: BLE001
            print("S=%d: flash_attn failed to run on this GPU: %r" % (S, e))
            return
        rel = [float((x.float() - y.float()).norm() / y.float().norm()) for x, y in zip(ro, rf)]
        flops = 3.5 * 4.0 * H * D * S * (S + 1) / 2
        t_o, t_f = timed(ours), timed(flash)
        print("S=%6d  |ours - flash_attn| / |flash_attn|  out %.2e dq %.2e dk %.2e dv %.2e" % (S, *rel))
        print("          lwm_b200 %.2f ms (%.0f TFLOP/s)   flash_attn %.2f ms (%.0f TFLOP/s)   speed-up x%.2f"
              % (t_o, flops / t_o / 1e9, t_f, flops / t_f / 1e9, t_f / t_o), flush=True)


if __name__ == "__main__":
    main()
