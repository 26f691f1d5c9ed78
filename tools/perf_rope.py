"""Throughput of the rotary-embedding kernel (lwm_attn_rope) at the 7B/128K shape: q and k [1,131072,32,128].
Algorithmic bytes = every element read once + written once."""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from lwm_b200 import _lib  # noqa: E402
from lwm_b200.rope import precompute_freqs_cis  # noqa: E402


def main():
    S, H = 131072, 32
    table = precompute_freqs_cis(128, 1 << 20, theta=5e7)
    pos = torch.arange(S, dtype=torch.int32, device="cuda")[None] + 900000
    for in_dt, out_dt in ((torch.bfloat16, torch.bfloat16), (torch.float32, torch.bfloat16), (torch.float32, torch.float32)):
        xq = torch.randn(1, S, H, 128, device="cuda").to(in_dt)
        xk = torch.randn(1, S, H, 128, device="cuda").to(in_dt)
        oq, ok = torch.empty_like(xq, dtype=out_dt), torch.empty_like(xk, dtype=out_dt)
        code = {torch.float32: 0, torch.bfloat16: 1}

        def run():
            _lib.call("lwm_attn_rope", _lib.ptr(xq), _lib.ptr(xk), code[in_dt], _lib.ptr(oq), _lib.ptr(ok), code[out_dt],
                      _lib.ptr(pos), _lib.ptr(table.inv_freq), 1, S, H, H, 128, 0, _lib.stream_ptr())
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        nbytes = 2 * S * H * 128 * (xq.element_size() + oq.element_size())
        print("rope %s -> %s: %.3f ms, %.0f GB/s (algorithmic %.2f GB)" % (in_dt, out_dt, ms, nbytes / ms / 1e6, nbytes / 1e9))
        del xq, xk, oq, ok


if __name__ == "__main__":
    main()
