"""VQGAN encode timing / profiling helper: python tools/perf_vqgan.py [frames] [precision]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lwm_b200.vqgan import VQGAN, init_params
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16x2"
params = init_params(seed=0)
x = (torch.rand(n, 256, 256, 3) * 2 - 1).cuda()
tok = VQGAN(params, precision=prec)
for _ in range(2):
    tok.encode(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    tok.encode(x)
b.record(); torch.cuda.synchronize()
print("encode %d frames (%s): %.2f ms" % (n, prec, a.elapsed_time(b) / 3))
# decode of the same number of frames (codes -> pixels; 477.4 GFLOP / frame)
codes = torch.randint(0, 8192, (n, 16, 16), dtype=torch.int32, device="cuda")
for _ in range(2):
    tok.decode(codes)
torch.cuda.synchronize()
a.record()
for _ in range(3):
    tok.decode(codes)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
print("decode %d frames (%s): %.2f ms = %.0f frames/s, %.0f algorithmic TFLOP/s" % (n, prec, ms, n / ms * 1e3,
                                                                                  n * 477.4e9 / ms / 1e9))
