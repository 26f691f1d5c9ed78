"""CPU validation (world 2 and 4, gloo + shared memory) of the EXPERIMENTAL one-sided ring executor
(lwm_b200/ring_exec_symm.py): K/V pulled from the owners' staged shards, dK/dV partials put into per-(step, sender,
block) landing slots and announced with per-step signals. The backend here emulates torch symmetric memory with
shared-memory CPU tensors (every rank can address every rank's workspace) and shared int32 mailboxes; numerics come from
the oracle-backed CPU step functions. Two passes are run back to back to exercise workspace / mailbox reuse."""
import contextlib
import os
import sys
import time

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_CHANNELS = 16


class EmulatedBackend:
    def __init__(self, rank, world, buffers, flags):
        self.rank, self.world, self.buffers, self.flags = rank, world, buffers, flags   # flags [world, channels, world]

    def reserve(self, nbytes):
        assert nbytes <= self.buffers[0].numel(), "test workspace too small"

    def view(self, peer, offset, shape, dtype):
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        return self.buffers[peer][offset:offset + n].view(dtype).view(*shape)

    def barrier(self):
        dist.barrier()

    def signal(self, peer, channel):
        assert int(self.flags[peer, channel, self.rank]) == 0, "mailbox still full: protocol error"
        self.flags[peer, channel, self.rank] = 1

    def wait_signal(self, peer, channel):
        t0 = time.time()
        while int(self.flags[self.rank, channel, peer]) == 0:
            assert time.time() - t0 < 60, "signal never arrived"
            time.sleep(0.001)
        self.flags[self.rank, channel, peer] = 0

    def on_pull_stream(self):
        return contextlib.nullcontext()

    on_put_stream = on_pull_stream

    def record(self, stream_name):
        return None

    def wait_event(self, stream_name, event):
        pass


def _worker(rank, world, port, layout, buffers, flags, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lwm_b200 import ring_exec_symm as rxs, ring_schedule as rs
        from oracle.step_ops import CpuOps
        from oracle.attn_dense import attention_dense, attention_dense_grads, finfo_min
        be = EmulatedBackend(rank, world, buffers, flags)
        B, S, H, D = 2, 256 * world, 2, 16
        Sl = S // world
        sl = slice(rank * Sl, (rank + 1) * Sl)
        errs = []
        for seed in (42, 43):                     # two passes: workspace and mailboxes are reused
            g = torch.Generator().manual_seed(seed)
            q, k, v, do = [torch.randn(B, S, H, D, generator=g) for _ in range(4)]
            bias = torch.zeros(B, S)
            bias[0, :37] = finfo_min("fp32")
            seg = torch.zeros(B, S, dtype=torch.int32)
            seg[1, S // 2 + 5:] = 1
            do[0, :37] = 0
            ks, vs = k[:, sl].contiguous(), v[:, sl].contiguous()
            plan = rs.make_plan(world, rank, Sl, Sl, True, layout)
            out, res = rxs.run_forward(plan, q[:, sl].contiguous(), ks, vs, bias, seg, True, None, CpuOps, be)
            dq, dk, dv = rxs.run_backward(plan, res, ks, vs, do[:, sl].contiguous(), bias, seg, True, None, CpuOps, be)
            kw = dict(causal=True, attn_bias=bias.numpy(), segment_ids=seg.numpy(), mask_value=finfo_min("fp32"))
            ref = attention_dense(q.numpy(), k.numpy(), v.numpy(), **kw)
            rq, rk, rv = attention_dense_grads(q.numpy(), k.numpy(), v.numpy(), do.numpy(), **kw)

            def err(x, r, skip_pad=False):
                x, r = x.double().numpy().copy(), r[:, sl].copy()
                if skip_pad and rank == 0:
                    x[0, :37], r[0, :37] = 0, 0       # padded query rows are arbitrary in the reference
                return float(np.linalg.norm(x - r) / np.linalg.norm(r))
            errs += [err(out, ref, True), err(dq, rq, True), err(dk, rk), err(dv, rv)]
        assert int(flags[rank].sum()) == 0            # every mailbox consumed
        ret[rank] = errs
    finally:
        dist.destroy_process_group()


def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,layout", [(2, "zigzag"), (4, "zigzag"), (4, "contiguous")])
def test_one_sided_executor_matches_dense_oracle(world, layout):
    buffers = [torch.zeros(8 << 20, dtype=torch.uint8).share_memory_() for _ in range(world)]
    flags = torch.zeros(world, N_CHANNELS, world, dtype=torch.int32).share_memory_()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), layout, buffers, flags, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert all(e < 1e-5 for e in ret[r]), (r, ret[r])


def test_landing_slots_are_consistent_between_sender_and_owner():
    sys.path.insert(0, ROOT)
    from lwm_b200 import ring_schedule as rs
    for P in (2, 4, 8):
        for layout in ("zigzag", "contiguous"):
            plans = [rs.make_plan(P, r, 1024, 1024, True, layout) for r in range(P)]
            for r, p in enumerate(plans):
                assert rs.peer_plan(plans[0], r).steps == p.steps
                table, rows = rs.landing_slots(p)
                spans = sorted((o, o + key[3]) for key, o in table.items())
                assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and (not spans or spans[-1][1] == rows)
            # every remote block a rank consumes has exactly one slot at its owner
            for r, p in enumerate(plans):
                for idx, st in enumerate(p.steps):
                    for kv in st.kv:
                        if kv.owner != r:
                            assert (idx, r, kv.start, kv.length) in rs.landing_slots(plans[kv.owner])[0]
