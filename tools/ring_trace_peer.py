"""Event timeline of ONE forward+backward of the sequence-parallel attention op through the peer-memory executor
(lwm_b200/ring_peer.py) at the headline shape, per rank: every labelled piece (scale exchange, staging, pulls, tile
kernels, partial puts, returns) with its start/end on its stream, relative to a common start barrier.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/ring_trace_peer.py [S_total]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    from lwm_b200 import ringattention as ra, ring_peer as rp, synthetic as syn
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    Sl, H, D = S // world, 32, 128
    q, k, v, do = [syn.shard(n_, rank, Sl, H, D).to(dev) for n_ in ("q", "k", "v", "do")]
    kw = dict(axis_name="sp", blockwise_kwargs=dict(causal_block_size=1))

    import time
    host = {}

    def step():
        qq, kk, vv = [t.detach().requires_grad_(True) for t in (q, k, v)]
        h0 = time.perf_counter()
        o = ra.ringattention(qq, kk, vv, None, None, **kw)
        h1 = time.perf_counter()
        o.backward(do)
        host["fwd_ms"], host["bwd_ms"] = (h1 - h0) * 1e3, (time.perf_counter() - h1) * 1e3
    for _ in range(3):
        step()
    tr = rp.CudaPeerTransport.get(dist.group.WORLD, dev)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    tr.trace = []
    t0 = torch.cuda.Event(enable_timing=True)
    t0.record()
    step()
    t1 = torch.cuda.Event(enable_timing=True)
    t1.record()
    torch.cuda.synchronize()
    spans = tr.span_times(t0, tr.trace)
    tr.trace = None
    total = t0.elapsed_time(t1)
    kern = sum(e - s for (lab, st, s, e) in spans if "kernel" in lab)
    lines = ["rank %d: pass %.2f ms, tile kernels %.2f ms (%.1f %%); host enqueue fwd %.2f ms, bwd %.2f ms" % (
        rank, total, kern, 100 * kern / total, host["fwd_ms"], host["bwd_ms"])]
    for (lab, st, s, e) in sorted(spans, key=lambda x: x[2]):
        lines.append("  %-5s %8.3f .. %8.3f  (%7.3f)  %s" % (st, s, e, e - s, lab))
    out = [None] * world
    dist.all_gather_object(out, "\n".join(lines))
    if rank == 0:
        for r in (0, world // 2, world - 1):
            print(out[r])
        print("summary: " + " | ".join(o.split("\n")[0] for o in out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
