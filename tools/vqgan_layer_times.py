"""Per-layer timing of the VQGAN conv stack (CUDA events): operand preparation and conv of the encoder's main layer
shapes in the 2-MMA fp16 scheme (with / without the GroupNorm-statistics epilogue) and the 3-MMA split-bf16 scheme.
  python tools/vqgan_layer_times.py [frames=16]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lwm_b200.vqgan import Ops, PackedConv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
g = torch.Generator().manual_seed(0)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3     # us


shapes = [(128, 128, 3, 256, 1), (128, 128, 3, 256, 2), (128, 256, 3, 128, 1), (256, 256, 3, 128, 1), (256, 256, 3, 64, 1),
          (256, 512, 3, 32, 1), (512, 512, 3, 32, 1), (512, 768, 3, 16, 1), (768, 768, 3, 16, 1), (128, 256, 1, 128, 1)]
o2, o3 = Ops("fp16x2"), Ops("bf16x3")
print("%-28s %9s %9s %9s %9s %9s %9s   %s" % ("layer", "prep16", "prep_hl", "conv2", "conv2+st", "conv3", "gnstats", "TFLOP/s issued conv2 / conv3"))
for cin, cout, k, H, stride in shapes:
    x = torch.randn(n, H, H, cin, generator=g).cuda()
    gn = {"scale": torch.ones(cin).cuda(), "bias": torch.zeros(cin).cuda()}
    pc = PackedConv({"kernel": torch.randn(k, k, cin, cout, generator=g) * 0.03, "bias": torch.zeros(cout)}, torch.device("cuda"))
    st = o2.gn_stats(x)
    x._gn_stats = st
    p16 = o2.prep(x, gn, n_pass=2)
    phl = o3.prep(x, gn, n_pass=3)
    t_p16 = timed(lambda: o2.prep(x, gn, n_pass=2))
    t_phl = timed(lambda: o3.prep(x, gn, n_pass=3))
    t_c2 = timed(lambda: o2.conv(p16, pc, stride=stride))
    t_c2s = timed(lambda: o2.conv(p16, pc, stride=stride, want_stats=True))
    t_c3 = timed(lambda: o3.conv(phl, pc, stride=stride))
    t_gs = timed(lambda: o2.gn_stats(x))
    Ho = H // stride
    fl = 2.0 * n * Ho * Ho * k * k * cin * cout
    print("%4d->%4d k%d %3dx%-3d s%d      %9.1f %9.1f %9.1f %9.1f %9.1f %9.1f   %6.0f / %6.0f" % (
        cin, cout, k, H, H, stride, t_p16, t_phl, t_c2, t_c2s, t_c3, t_gs, 2 * fl / t_c2 / 1e6, 3 * fl / t_c3 / 1e6))
