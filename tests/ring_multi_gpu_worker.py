"""torchrun worker: ring attention on N GPUs vs the single-GPU result of the same kernels.
Launched by tests/test_ring_multi_gpu.py (and by hand:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/ring_multi_gpu_worker.py)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from lwm_b200.ringattention import ringattention
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    S = int(os.environ.get("RING_TEST_S", str(1024 * world)))
    B, H, D = 1, 4, 128
    g = torch.Generator().manual_seed(99)
    q, k, v, do = [torch.randn(B, S, H, D, generator=g).to(torch.bfloat16).to(dev) for _ in range(4)]
    bias = torch.zeros(B, 1, 1, S)
    bias[..., :50] = -3.3895313892515355e38
    seg = torch.zeros(B, S, dtype=torch.int32)
    seg[:, S // 2 + 7:] = 1
    kw = dict(axis_name="sp", float32_logits=True, cache_idx=None,
              blockwise_kwargs=dict(causal_block_size=1, deterministic=True, attn_pdrop=0.0, query_chunk_size=256,
                                    key_chunk_size=256))
    Sl = S // world
    sl = slice(rank * Sl, (rank + 1) * Sl)
    worst = 0.0
    for layout in ("contiguous", "zigzag"):
        for masks in (False, True):
            b_, s_ = (bias.to(dev), seg.to(dev)) if masks else (None, None)
            dol = do.clone()
            if masks:
                dol[:, :50] = 0
            # single-GPU result of the same op (ring size 1) as the comparison point
            import lwm_b200.ringattention as ra
            qf, kf, vf = [t.detach().clone().requires_grad_(True) for t in (q, k, v)]
            saved = ra._resolve_group
            ra._resolve_group = lambda axis: (None, 0, 1)
            of = ringattention(qf, kf, vf, b_, s_, **kw)
            of.backward(dol)
            ra._resolve_group = saved
            ql, kl, vl = [t[:, sl].detach().clone().requires_grad_(True) for t in (q, k, v)]
            ol = ringattention(ql, kl, vl, b_, s_, layout=layout, **kw)
            ol.backward(dol[:, sl].contiguous())
            torch.cuda.synchronize()

            def rel(a, b2):
                a, b2 = a.float(), b2.float()
                if masks and rank == 0:
                    a, b2 = a[:, 50:], b2[:, 50:]
                return float((a - b2).norm() / b2.norm().clamp_min(1e-30))
            errs = (rel(ol, of[:, sl]), rel(ql.grad, qf.grad[:, sl]), rel(kl.grad, kf.grad[:, sl]),
                    rel(vl.grad, vf.grad[:, sl]))
            worst = max(worst, *errs)
            print("rank %d layout=%s masks=%s errs(out,dq,dk,dv)=%s" % (rank, layout, masks,
                                                                         ["%.2e" % e for e in errs]), flush=True)
    # decode path: every rank holds a KV-cache shard, the single query row is replicated
    from lwm_b200.ringattention import ringattention_inference
    g2 = torch.Generator().manual_seed(7)
    qd = torch.randn(B, 1, H, D, generator=g2).to(torch.bfloat16).to(dev)
    maskd = torch.ones(B, 1, 1, S, dtype=torch.bool)
    maskd[..., :29] = False
    maskd[..., S - 5:] = False
    od = ringattention_inference(qd, k[:, sl].contiguous(), v[:, sl].contiguous(), maskd.to(dev), axis_name="sp")
    saved = ra._resolve_group
    ra._resolve_group = lambda axis: (None, 0, 1)
    od_ref = ringattention_inference(qd, k, v, maskd.to(dev), axis_name="sp")
    ra._resolve_group = saved
    torch.cuda.synchronize()
    e = float((od.float() - od_ref.float()).norm() / od_ref.float().norm())
    print("rank %d decode (ringattention_inference) err vs single-GPU = %.2e" % (rank, e), flush=True)
    worst = max(worst, e)
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.destroy_process_group()
    # identical kernels, different tiling of the work: only bf16 re-rounding of out / grads differs
    if t.item() > 1e-2:
        print("RING_MULTI_GPU FAIL worst=%.3e" % t.item())
        sys.exit(1)
    if rank == 0:
        print("RING_MULTI_GPU OK worst=%.3e" % t.item())


if __name__ == "__main__":
    main()
