"""CPU validation of the peer-memory executor (lwm_b200/ring_peer.py) against the dense fp64 oracle: P rank threads,
emulated heaps/flags (tests/peer_emulation.py), oracle-backed step functions. Covers both work assignments, padding
bias + packed segments, the scaled (fp16-mode) and unscaled operand bookkeeping, fp32 and bf16 results, B > 1, and
several passes back to back (heap-region reuse across pass parity)."""
import os
import sys
import threading

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _run_world(world, layout, causal, scaled, want_f32, B=1, Sl=256, H=2, D=16, passes=2, masks=True):
    from lwm_b200 import ring_peer as rp, ring_schedule as rs
    from oracle.attn_dense import attention_dense, attention_dense_grads, finfo_min
    from peer_emulation import EmuWorld, EmuTransport, EmuOps
    emu = EmuWorld(world)
    S = Sl * world
    errs, fails, logs = {}, [], {}

    def worker(rank):
        try:
            tr = EmuTransport(emu, rank)
            ops = EmuOps(scaled)
            sl = slice(rank * Sl, (rank + 1) * Sl)
            plan = rs.make_peer_plan(world, rank, Sl, Sl, causal, layout, fwd_group_chunks=2)
            out_errs = []
            for seed in range(passes):
                g = torch.Generator().manual_seed(100 + seed)
                q, k, v, do = [torch.randn(B, S, H, D, generator=g) * (0.5 + seed) for _ in range(4)]
                bias = seg = None
                npad = 0
                if masks:
                    npad = 37
                    bias = torch.zeros(B, S)
                    bias[0, :npad] = finfo_min("fp32")
                    seg = torch.zeros(B, S, dtype=torch.int32)
                    seg[B - 1, S // 2 + 5:] = 1
                    do[0, :npad] = 0
                in_dtype = torch.float32 if want_f32 else torch.bfloat16
                if not want_f32:
                    q, k, v, do = [t.to(torch.bfloat16).float() for t in (q, k, v, do)]
                ql, kl, vl, dl = [t[:, sl].contiguous().to(in_dtype) for t in (q, k, v, do)]
                out, res = rp.run_forward(plan, ql, kl, vl, bias, seg, causal, ops, tr, want_f32)
                dq, dk, dv = rp.run_backward(plan, res, kl, vl, dl, bias, seg, causal, ops, tr, want_f32)
                assert out.dtype == in_dtype and dq.dtype == in_dtype and dk.dtype == in_dtype
                kw = dict(causal=causal, mask_value=finfo_min("fp32"))
                if masks:
                    kw.update(attn_bias=bias.numpy(), segment_ids=seg.numpy())
                ref = attention_dense(q.numpy(), k.numpy(), v.numpy(), **kw)
                rq, rk, rv = attention_dense_grads(q.numpy(), k.numpy(), v.numpy(), do.numpy(), **kw)

                def err(x, r, skip_pad=False):
                    x, r = x.double().numpy().copy(), r[:, sl].copy()
                    if skip_pad and rank == 0 and npad:
                        x[0, :npad], r[0, :npad] = 0, 0
                    return float(np.linalg.norm(x - r) / np.linalg.norm(r))
                out_errs += [err(out, ref, True), err(dq, rq, True), err(dk, rk), err(dv, rv)]
            errs[rank] = out_errs
            logs[rank] = tr.log
        except BaseException as e:   # noqa: BLE001  (propagated to the main thread)
            import traceback
            fails.append((rank, traceback.format_exc()))
            emu.barrier.abort()

    ts = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert not fails, fails[0][1]
    assert len(errs) == world
    return errs, logs


@pytest.mark.parametrize("world,layout,causal", [(2, "zigzag", True), (4, "zigzag", True), (8, "zigzag", True),
                                                 (4, "contiguous", True), (3, "contiguous", False)])
def test_peer_executor_matches_dense_oracle_fp32_results(world, layout, causal):
    errs, _ = _run_world(world, layout, causal, scaled=True, want_f32=True)
    for r in range(world):
        assert all(e < 2e-5 for e in errs[r]), (r, errs[r])


@pytest.mark.parametrize("world,scaled", [(2, False), (4, True)])
def test_peer_executor_bf16_results_and_batch(world, scaled):
    errs, _ = _run_world(world, "zigzag", True, scaled=scaled, want_f32=False, B=2, passes=3)
    for r in range(world):
        assert all(e < 6e-3 for e in errs[r]), (r, errs[r])     # bf16 rounding of out / grads only


def test_traffic_matches_the_plan():
    """every rank pulls exactly the K/V chunks (+ its remote Q/dO chunks) the plan lists and nothing else"""
    from lwm_b200 import ring_schedule as rs
    world, Sl, H, D = 4, 256, 2, 16
    _, logs = _run_world(world, "zigzag", True, scaled=False, want_f32=True, Sl=Sl, H=H, D=D, passes=1, masks=False)
    row = H * D * 4
    for r in range(world):
        plan = rs.make_peer_plan(world, r, Sl, Sl, True, "zigzag")
        kv_rows = sum(c.length for g in plan.bwd_groups for c in g.chunks if c.owner != r)
        q_rows = sum(qc.length for qc in plan.q_chunks if qc.owner != r)
        pulled = sum(n for (kind, _, n) in logs[r] if kind == "pull")
        assert pulled == 2 * (2 * kv_rows + q_rows) * row          # fwd + bwd, K and V
        put = sum(n for (kind, _, n) in logs[r] if kind == "put")
        # fwd: O chunks (fp32); bwd: dQ chunks (fp32) + dK/dV partial per remote chunk (fp32)
        assert put == (2 * q_rows + 2 * kv_rows) * row


def test_plan_properties():
    from lwm_b200 import ring_schedule as rs
    for P in (2, 4, 8):
        plans = [rs.make_peer_plan(P, r, 1024, 1024, True) for r in range(P)]
        for r, p in enumerate(plans):
            # balanced causal work: every rank computes the same number of (q chunk, kv chunk) tile-equivalents
            work = 0.0
            for g in p.bwd_groups:
                c = g.chunks[0]
                for (qi, p0, rows, _o) in g.launches:
                    work += 0.5 if p.q_chunks[qi].pos0 == p0 else 1.0
            assert work == 2.0 * P, (P, r, work)
            # every partial a rank sends has a landing slot at its owner, and vice versa
            for g in p.bwd_groups:
                c = g.chunks[0]
                if c.owner != r:
                    assert (c.index, r) in plans[c.owner].incoming
            for (ci, peer) in p.incoming:
                assert any(g.chunks[0].owner == r and g.chunks[0].index == ci for g in plans[peer].bwd_groups)
            # forward and backward visit the same set of chunks
            f = sorted((c.owner, c.index) for g in p.fwd_groups for c in g.chunks)
            b = sorted((c.owner, c.index) for g in p.bwd_groups for c in g.chunks)
            assert f == b


def test_flag_ranges_do_not_collide():
    """STAGED[rank], RES[rank] and PART[slot] live in one flag page: their index ranges are disjoint up to the largest
    supported ring (16 ranks, 2 chunks per rank) and stay inside the page"""
    from lwm_b200 import ring_peer as rp, ring_schedule as rs
    world = 16
    staged = set(range(rp.FLAG_STAGED, rp.FLAG_STAGED + world))
    res = set(range(rp.FLAG_RES, rp.FLAG_RES + world))
    plan = rs.make_peer_plan(world, 0, 1024, 1024, True, "zigzag")
    part = {rp.FLAG_PART + plan.slot(ci, peer) for ci in range(plan.chunks_per_rank) for peer in range(world)}
    assert not (staged & res) and not (staged & part) and not (res & part)
    assert max(part) < 65536 // 4
    lay = rp.Layout(1, 1024, 1024, 32, 128, world, 2, 2)
    assert lay.n_slots == max(plan.slot(ci, peer) for ci in range(2) for peer in range(world)) + 1
