"""Timeline of one forward+backward ring pass (debug): torchrun --nproc-per-node N tools/ring_trace.py [S_total]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from lwm_b200 import ring_exec as rx, ringattention as ra

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
dist.init_process_group("nccl", device_id=dev)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
Sl = S // world
q, k, v, do = [torch.randn(1, Sl, 32, 128, device=dev).to(torch.bfloat16) for _ in range(4)]


def one():
    qq, kk, vv = [t.detach().requires_grad_(True) for t in (q, k, v)]
    out = ra.ringattention(qq, kk, vv, None, None, blockwise_kwargs=dict(causal_block_size=1))
    out.backward(do)


for _ in range(3):
    one()
torch.cuda.synchronize(); dist.barrier()
hp = rx._HP_GROUPS
if rank == 0:
    print("hp groups:", {k: (type(g).__name__, g is not None and g is not dist.group.WORLD) for k, g in hp.items()}, flush=True)
rx.trace_begin()
t0 = torch.cuda.Event(enable_timing=True); t0.record()
one()
spans = rx.trace_end(t0)
t1 = max(e for _, _, _, e in spans)
if rank in (0, world - 1):
    print("rank %d: pass %.2f ms" % (rank, t1), flush=True)
    for lab, st, a, b in sorted(spans, key=lambda x: x[2]):
        print("  r%d %-5s %8.2f -> %8.2f (%6.2f ms)  %s" % (rank, st, a, b, b - a, lab), flush=True)
    busy = sum(b - a for lab, st, a, b in spans if st == "main")
    print("  r%d main-stream kernel time %.2f ms of %.2f ms" % (rank, busy, t1), flush=True)
dist.barrier(); dist.destroy_process_group()
