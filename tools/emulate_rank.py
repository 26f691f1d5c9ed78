"""Run ONE rank's share of a P-way ring plan on a single GPU with the NCCL exchange stubbed out
(received blocks are just local buffers). Separates host/launch/kernel-granularity overhead from
communication effects when analysing multi-GPU scaling. usage: emulate_rank.py [P] [rank] [S_total]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lwm_b200 import ring_exec as rx, ring_schedule as rs, ringattention as ra

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 3
S = int(sys.argv[3]) if len(sys.argv) > 3 else 131072
Sl = S // P
B, H, D = 1, 32, 128


class FakeComm(rx._Comm):
    def __init__(self, group, device, channel=0):
        self.group, self.device, self.cuda, self.stream = group, device, False, None

    def exchange(self, sends, recvs, after_event=None):
        return None

    def wait(self, token):
        return


rx._Comm = FakeComm
q, k, v, do = [torch.randn(B, Sl, H, D, device="cuda").to(torch.bfloat16) for _ in range(4)]
plan = rs.make_plan(P, rank, Sl, Sl, True, "zigzag")
n_pairs = sum(len(st.pairs) for st in plan.steps)
print("P=%d rank=%d S_loc=%d: %d steps, %d (q chunk, kv block) launches per pass" % (P, rank, Sl, len(plan.steps), n_pairs))


def one():
    out, res = rx.run_forward(plan, q, k, v, None, None, True, None, ra.CudaOps)
    return rx.run_backward(plan, res, k, v, do, None, None, True, None, ra.CudaOps)


for _ in range(2):
    one()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 4
t0 = time.perf_counter()
e0.record()
for _ in range(n):
    one()
e1.record()
t_cpu = (time.perf_counter() - t0) / n
torch.cuda.synchronize()
gpu_ms = e0.elapsed_time(e1) / n
f = 3.5 * 4.0 * H * D * S * (S + 1) / 2 / P
print("per pass (fwd+bwd): GPU %.2f ms, host enqueue %.2f ms; algorithmic %.1f TFLOP/s" % (gpu_ms, t_cpu * 1e3, f / gpu_ms / 1e9))
# kernel-only time: the same launches timed individually
import collections
times = collections.defaultdict(float)
orig_f, orig_b = ra.CudaOps.fwd_step, ra.CudaOps.bwd_step


def timed(name, fn):
    def w(*a, **kw):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(); fn(*a, **kw); a1.record(); torch.cuda.synchronize()
        times[name] += a0.elapsed_time(a1)
    return staticmethod(w)


ra.CudaOps.fwd_step = timed("fwd_kernels", orig_f)
ra.CudaOps.bwd_step = timed("bwd_kernels", orig_b)
one()
print("sum of tile-kernel times in one pass: fwd %.2f ms, bwd %.2f ms, total %.2f ms" % (times["fwd_kernels"], times["bwd_kernels"], sum(times.values())))
