"""Quick device-side timing of the attention tile kernels (development aid, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lwm_b200 import ringattention as ra


def time_fn(fn, warm=2, iters=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
    for S in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["4096", "16384", "32768"])]:
        B, H, D = 1, 32, 128
        q, k, v, do = [torch.randn(B, S, H, D, device="cuda").to(torch.bfloat16) for _ in range(4)]
        out = torch.empty_like(q)
        fsc = bsc = None
        if (sys.argv[3] if len(sys.argv) > 3 else "fp16") == "fp16":     # the default precision mode's kernels
            (q, sq), (k, sk), (v, sv), (do16, sd) = [ra.to_f16(t) for t in (q, k, v, do)]
            fsc, bsc = (sq, sk, sv), (sq, sk, sv, sd)
        lse = torch.empty(B, H, S, dtype=torch.float32, device="cuda")
        f_fwd = 4.0 * B * H * D * S * (S + 1) / 2
        ms = time_fn(lambda: ra.fwd_step(q, k, v, out, lse, None, None, None, 0, 0, True, None, None, True, True, scales=fsc))
        print("fwd  causal S=%6d: %8.3f ms  %7.1f TFLOP/s" % (S, ms, f_fwd / ms / 1e9))
        if which == "bwd":
            delta = torch.empty_like(lse)
            dq = torch.zeros(B, S, H, D, dtype=torch.float32, device="cuda")
            dk = torch.zeros_like(dq)
            dv = torch.zeros_like(dq)
            ra.bwd_prep(out, do, delta)
            nl = ra.lse_for_bwd(lse, f16=bsc is not None)
            dd = do16 if bsc is not None else do
            ms = time_fn(lambda: ra.bwd_step(q, k, v, dd, nl, delta, dq, dk, dv, 0, 0, True, None, None, scales=bsc))
            print("bwd  causal S=%6d: %8.3f ms  %7.1f TFLOP/s" % (S, ms, 2.5 * f_fwd / ms / 1e9))


if __name__ == "__main__":
    main()
