"""VQGAN encode timing / profiling helper: python tools/perf_vqgan.py [frames] [precision]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lwm_b200.vqgan import VQGAN, init_params
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
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
