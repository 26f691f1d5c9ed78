"""World-size-2/4 CPU tests (gloo) of the ring sequencing code: the same lwm_b200.ring_exec that
drives the CUDA kernels, with the oracle-backed CPU step functions injected. Checks, against the
dense float64 oracle on the full (un-sharded) sequence:
  * forward output and backward dq/dk/dv of every rank's contiguous shard,
  * for both schedules (reference 'contiguous' and load-balanced 'zigzag'),
  * with a left-padding bias + segment ids (global-position indexing through the permutation)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, layout, use_masks, ret, n_sub=1, batch=1, prefetch="1"):
    sys.path.insert(0, ROOT)
    os.environ["LWM_RING_PREFETCH"] = prefetch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lwm_b200 import ring_exec as rx, ring_schedule as rs
        from oracle.step_ops import CpuOps
        from oracle.attn_dense import attention_dense, attention_dense_grads, finfo_min
        torch.manual_seed(0)
        B, S, H, D = batch, 256 * world * n_sub, 2, 16   # zigzag half-chunks of 128 * n_sub rows
        Sl = S // world
        g = torch.Generator().manual_seed(42)
        q, k, v, do = [torch.randn(B, S, H, D, generator=g) for _ in range(4)]
        bias = seg = None
        if use_masks:
            bias = torch.zeros(B, S)
            bias[0, :37] = finfo_min("fp32")
            seg = torch.zeros(B, S, dtype=torch.int32)
            seg[0, S // 2 + 5:] = 1
            do = do.clone()
            do[:, :37] = 0
        sl = slice(rank * Sl, (rank + 1) * Sl)
        plan = rs.make_plan(world, rank, Sl, Sl, True, layout, n_sub_first=n_sub)
        out, res = rx.run_forward(plan, q[:, sl].contiguous(), k[:, sl].contiguous(), v[:, sl].contiguous(), bias,
                                  seg, True, None, CpuOps)
        plan = rs.make_plan(world, rank, Sl, Sl, True, layout, n_sub_first=n_sub, n_sub_last=n_sub)
        dq, dk, dv = rx.run_backward(plan, res, k[:, sl].contiguous(), v[:, sl].contiguous(),
                                     do[:, sl].contiguous(), bias, seg, True, None, CpuOps)
        kw = dict(causal=True, attn_bias=None if bias is None else bias.numpy(),
                  segment_ids=None if seg is None else seg.numpy(), mask_value=finfo_min("fp32"))
        ref = attention_dense(q.numpy(), k.numpy(), v.numpy(), **kw)
        rq, rk, rv = attention_dense_grads(q.numpy(), k.numpy(), v.numpy(), do.numpy(), **kw)
        lo = 37 if use_masks else 0          # padded query rows are arbitrary in the reference

        def err(x, r):
            x = x.double().numpy()
            r = r[:, sl]
            if rank == 0 and lo:
                x, r = x[:, lo:], r[:, lo:]
            return float(np.linalg.norm(x - r) / max(np.linalg.norm(r), 1e-30))
        ret[rank] = (err(out, ref), err(dq, rq), float(np.linalg.norm(dk.double().numpy() - rk[:, sl]) /
                                                       np.linalg.norm(rk[:, sl])),
                     float(np.linalg.norm(dv.double().numpy() - rv[:, sl]) / np.linalg.norm(rv[:, sl])))
    finally:
        dist.destroy_process_group()


def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,layout,use_masks", [(2, "contiguous", False), (2, "zigzag", False),
                                                    (2, "zigzag", True), (4, "zigzag", False),
                                                    (4, "contiguous", True)])
def test_ring_schedules_match_dense_oracle(world, layout, use_masks):
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), layout, use_masks, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        for e in ret[r]:
            assert e < 1e-5, (r, ret[r])


def test_ring_batch_gt_one_matches_dense_oracle():
    """B > 1: sequence slices of [B,S,H,D] are no longer contiguous views (staged copies in ring_exec)"""
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, _free_port(), "zigzag", True, ret, 1, 2), nprocs=2, join=True)
    for r in range(2):
        for e in ret[r]:
            assert e < 1e-5, (r, ret[r])


@pytest.mark.parametrize("world,layout", [(2, "zigzag"), (4, "zigzag"), (2, "contiguous")])
def test_sub_step_pipelined_plans_match_dense_oracle(world, layout):
    """first/last step cut in 2 sub-steps (transfer of piece j+1 overlaps the kernels of piece j)"""
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), layout, True, ret, 2), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        for e in ret[r]:
            assert e < 1e-5, (r, ret[r])


@pytest.mark.parametrize("world,layout,prefetch", [(4, "zigzag", "all"), (4, "contiguous", "2")])
def test_deeper_kv_prefetch_matches_dense_oracle(world, layout, prefetch):
    """LWM_RING_PREFETCH: every (or several) K/V exchange(s) posted ahead instead of one step ahead"""
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), layout, True, ret, 1, 1, prefetch), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        for e in ret[r]:
            assert e < 1e-5, (r, ret[r])


def test_zigzag_plan_is_balanced_and_consistent():
    sys.path.insert(0, ROOT)
    from lwm_b200 import ring_schedule as rs
    for P in (2, 4, 8):
        plans = [rs.make_plan(P, r, 1024, 1024, True, "zigzag") for r in range(P)]
        per_step = [rs.work_units(p, True) for p in plans]
        flat = [w for ws in per_step for w in ws]
        assert max(flat) == min(flat)                       # identical causal work, every rank, every step
        # every send has exactly one matching receive
        for n_sub in (1, 2):
            pl = plans if n_sub == 1 else [rs.make_plan(P, r, 1024, 1024, True, "zigzag", 2, 2) for r in range(P)]
            assert len({len(p.steps) for p in pl}) == 1
            for idx in range(len(pl[0].steps)):
                sends = sorted((r, peer, s, l) for r, p in enumerate(pl) for (s, l, peer) in p.steps[idx].sends)
                recvs = sorted((kv.owner, r, kv.start, kv.length) for r, p in enumerate(pl)
                               for kv in p.steps[idx].kv if kv.owner != r)
                assert sends == recvs
        contiguous = [sum(rs.work_units(rs.make_plan(P, r, 1024, 1024, True, "contiguous"), True)) for r in range(P)]
        assert max(contiguous) / (sum(contiguous) / P) > 1.4   # the imbalance zigzag removes


def _worker_f16_plumbing(rank, world, port, ret):
    """The fp16-precision ops wrapper (operand-conversion cache keyed by buffer address, fp32 output residuals) driven
    through the real ring sequencing under gloo, with its three kernel entry points replaced by CPU emulations."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lwm_b200 import ring_exec as rx, ring_schedule as rs, ringattention as ra
        from oracle.step_ops import CpuOps
        from oracle.attn_dense import attention_dense, attention_dense_grads, finfo_min

        def to_f16(x):
            e = torch.floor(torch.log2(x.abs().max()))
            scale = torch.pow(torch.tensor(2.0), e - 12)
            return (x / scale).half(), scale

        def fwd_step(q16, k16, v16, out, lse, acc_o, acc_m, acc_l, q_pos0, k_pos0, causal, bias, seg, first, last,
                     scales=None, out_f32=None):
            sq, sk, sv = scales
            CpuOps.fwd_step(q16.float() * sq, k16.float() * sk, v16.float() * sv, out, lse, acc_o, acc_m, acc_l,
                            q_pos0, k_pos0, causal, bias, seg, first, last)
            if out_f32 is not None:
                out_f32.copy_(out)

        def bwd_step(q16, k16, v16, d16, lse, delta, dq, dk, dv, q_pos0, k_pos0, causal, bias, seg, scales=None):
            sq, sk, sv, sd = scales
            CpuOps.bwd_step(q16.float() * sq, k16.float() * sk, v16.float() * sv, d16.float() * sd, lse, delta, dq, dk,
                            dv, q_pos0, k_pos0, causal, bias, seg)
        ra.to_f16, ra.fwd_step, ra.bwd_step = to_f16, fwd_step, bwd_step

        class Ops(ra.CudaOpsF16):
            bwd_prep = staticmethod(CpuOps.bwd_prep)
            lse_for_bwd = staticmethod(CpuOps.lse_for_bwd)
            cast = staticmethod(CpuOps.cast)
            accumulate = staticmethod(CpuOps.accumulate)

        B, S, H, D = 1, 256 * world, 2, 16
        Sl = S // world
        g = torch.Generator().manual_seed(7)
        q, k, v, do = [torch.randn(B, S, H, D, generator=g) for _ in range(4)]
        sl = slice(rank * Sl, (rank + 1) * Sl)
        ks, vs = k[:, sl].contiguous(), v[:, sl].contiguous()
        plan = rs.make_plan(world, rank, Sl, Sl, True, "zigzag")
        ops = Ops()
        out, res = rx.run_forward(plan, q[:, sl].contiguous(), ks, vs, None, None, True, None, ops)
        n_cached = len(ops._cache)
        res = ra._f32_residuals(ops, res)
        assert all(o.dtype == torch.float32 for o in res["out_chunks"]) and len(ops.out_f32) == len(res["out_chunks"])
        dq, dk, dv = rx.run_backward(plan, res, ks, vs, do[:, sl].contiguous(), None, None, True, None, Ops())
        kw = dict(causal=True, attn_bias=None, segment_ids=None, mask_value=finfo_min("fp32"))
        ref = attention_dense(q.numpy(), k.numpy(), v.numpy(), **kw)
        rq, rk, rv = attention_dense_grads(q.numpy(), k.numpy(), v.numpy(), do.numpy(), **kw)

        def err(x, r):
            return float(np.linalg.norm(x.double().numpy() - r[:, sl]) / np.linalg.norm(r[:, sl]))
        ret[rank] = (err(out, ref), err(dq, rq), err(dk, rk), err(dv, rv), n_cached,
                     len(plan.q_chunks) + 2 * sum(len(st.kv) for st in plan.steps))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_fp16_precision_wrapper_through_the_ring(world):
    ret = mp.Manager().dict()
    mp.spawn(_worker_f16_plumbing, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert all(e < 2e-3 for e in ret[r][:4]), (r, ret[r])      # fp16-rounded operands: ~3e-4; a stale cache hit: O(1)
        # every distinct operand block converted exactly once: the q chunks + K and V of every visible block
        assert ret[r][4] == ret[r][5], ret[r]
