"""EXPERIMENTAL one-sided executor of a ring Plan (opt-in: LWM_RING_TRANSPORT=symm) — DESIGN.md §7, round-2 item 1.

Status: the sequencing and the slot / signal protocol below are validated on CPU with an emulated backend
(tests/test_ring_symm_emulated.py: shared-memory buffers + flags across gloo ranks); the CUDA backend
(`SymmMemBackend`, torch symmetric memory = CUDA VMM allocations mapped into every peer over NVLink) has NOT yet run
on hardware. The default transport remains ring_exec.py (NCCL send/recv), which is the measured one.

Why: with two-sided send/recv every transfer needs both ranks' communication streams to reach the matching call; at
8 GPUs the measured timelines (profiles/ring_timeline_n8_r01.log) show late peers head-of-line blocking the K/V
prefetch of later steps and a ~7 ms tail of dK/dV returns. K/V are immutable during a pass, so nothing has to be
negotiated:
  * every rank stages its K/V shard once in a peer-mapped buffer (one barrier per pass);
  * a rank that needs a block PULLS it from the owner's buffer with a plain device-to-device copy on a side stream
    (copy engines, no SMs, no action by the owner); all pulls of a pass are posted up front in step order;
  * dK/dV partials are PUT into a per-(step, sender, block) slot of the owner's landing zone, followed by a signal;
    the owner folds the slots into its accumulators after its own last tile kernel, waiting on the signals in order.
Q / dO / O / dQ permutations of the zigzag layout keep using ring_exec's helpers (small, one exchange per pass).

Workspace layout (identical on every rank, byte offsets):  [ K shard | V shard | landing dK rows | landing dV rows ].
Signal channels: 0 = barriers, 1 + step = "partials of that step are in your landing zone".
"""
import torch
import torch.distributed as dist

from . import ring_exec as rx
from . import ring_schedule as rs


class SymmMemBackend:
    """torch.distributed._symmetric_memory backend (one instance per (group, device), workspace grows on demand)."""

    _instances = {}

    @classmethod
    def get(cls, group, device):
        key = (id(group) if group is not None else 0, device.index)
        if key not in cls._instances:
            cls._instances[key] = cls(group, device)
        return cls._instances[key]

    def __init__(self, group, device):
        self.group = group if group is not None else dist.group.WORLD
        self.device = device
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.nbytes = 0
        self.buf = self.hdl = None
        self.pull_stream = torch.cuda.Stream(device=device, priority=-1)
        self.put_stream = torch.cuda.Stream(device=device, priority=-1)

    def reserve(self, nbytes):
        """collective: every rank calls it with the same size"""
        if nbytes <= self.nbytes:
            return
        import torch.distributed._symmetric_memory as symm
        nbytes = (nbytes + (1 << 21) - 1) >> 21 << 21
        symm.enable_symm_mem_for_group(self.group.group_name)       # idempotent: registers the group's store
        self.buf = symm.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.hdl = symm.rendezvous(self.buf, self.group)
        self.nbytes = nbytes

    def view(self, peer, offset, shape, dtype):
        """tensor view of `peer`'s workspace at byte `offset` (peer == my rank: local memory)"""
        itemsize = torch.empty((), dtype=dtype).element_size()
        assert offset % itemsize == 0
        return self.hdl.get_buffer(peer, tuple(shape), dtype, offset // itemsize)

    def barrier(self):
        self.hdl.barrier(channel=0)

    def signal(self, peer, channel):
        self.hdl.put_signal(peer, channel=channel)

    def wait_signal(self, peer, channel):
        self.hdl.wait_signal(peer, channel=channel)

    # stream plumbing (the emulated CPU backend makes all of these no-ops)
    def on_pull_stream(self):
        return torch.cuda.stream(self.pull_stream)

    def on_put_stream(self):
        return torch.cuda.stream(self.put_stream)

    def record(self, stream_name):
        s = {"main": torch.cuda.current_stream(self.device), "pull": self.pull_stream, "put": self.put_stream}[stream_name]
        return s.record_event()

    def wait_event(self, stream_name, event):
        s = {"main": torch.cuda.current_stream(self.device), "pull": self.pull_stream, "put": self.put_stream}[stream_name]
        s.wait_event(event)


def _layout(plan, k, v):
    """byte offsets of the workspace regions for this call"""
    kv_bytes = k.numel() * k.element_size()
    _, rows = rs.landing_slots(plan)
    # every rank reserves for the largest landing zone of the ring (plans differ per rank; the workspace is symmetric)
    max_rows = max(rs.landing_slots(rs.peer_plan(plan, r))[1] for r in range(plan.world))
    B, _, H, D = k.shape
    land_bytes = B * max_rows * H * D * 4
    off = dict(k=0, v=kv_bytes, dk=2 * kv_bytes, dv=2 * kv_bytes + land_bytes)
    return off, 2 * kv_bytes + 2 * land_bytes, max_rows


def _stage_kv(be, off, k, v):
    """copy my K/V shard into the peer-visible workspace; after the barrier every rank may pull from every rank"""
    ev = be.record("main")
    be.wait_event("pull", ev)
    with be.on_pull_stream():
        be.view(be.rank, off["k"], k.shape, k.dtype).copy_(k)
        be.view(be.rank, off["v"], v.shape, v.dtype).copy_(v)
        be.barrier()


def _post_pulls(be, plan, off, k, v):
    """post the pulls of EVERY step up front (K/V are immutable); -> per step (bufs, event).
    The receive buffers come from the MAIN stream's allocator pool (they are consumed there); the pull stream first
    waits for everything already queued on the main stream, so a recycled block cannot still be in use."""
    be.wait_event("pull", be.record("main"))
    out = []
    for st in plan.steps:
        bufs, copies = [], []
        for kv in st.kv:
            if kv.owner == plan.rank:
                bufs.append((rx._rows(k, kv.start, kv.length), rx._rows(v, kv.start, kv.length)))
                continue
            kb = torch.empty((k.shape[0], kv.length) + tuple(k.shape[2:]), dtype=k.dtype, device=k.device)
            vb = torch.empty_like(kb)
            bufs.append((kb, vb))
            copies.append((kb, vb, kv))
        with be.on_pull_stream():
            for kb, vb, kv in copies:
                kb.copy_(be.view(kv.owner, off["k"], k.shape, k.dtype)[:, kv.start:kv.start + kv.length])
                vb.copy_(be.view(kv.owner, off["v"], v.shape, v.dtype)[:, kv.start:kv.start + kv.length])
            out.append((bufs, be.record("pull")))
    return out


def _end_pass(be):
    """nobody may re-stage its K/V (next pass) before every rank has finished pulling: barrier on the pull stream"""
    with be.on_pull_stream():
        be.barrier()


def run_forward(plan, q, k, v, bias, seg, causal, group, ops, be):
    dev = q.device
    comm = rx._Comm(group, dev)
    B, Sq, H, D = q.shape
    off, nbytes, _ = _layout(plan, k, v)
    be.reserve(nbytes)
    _stage_kv(be, off, k, v)
    q_chunks = rx._gather_q_like(plan, comm, q)
    n_q = len(q_chunks)
    out_chunks = [torch.empty_like(c) for c in q_chunks]
    lse_chunks = [torch.empty((B, H, c.shape[1]), dtype=torch.float32, device=dev) for c in q_chunks]
    visits = [[] for _ in range(n_q)]
    for idx, st in enumerate(plan.steps):
        for (qi, ki) in st.pairs:
            visits[qi].append((idx, ki))
    acc = [None] * n_q
    for qi in range(n_q):
        if len(visits[qi]) > 1:
            c = q_chunks[qi]
            acc[qi] = (torch.empty(c.shape, dtype=torch.float32, device=dev),
                       torch.empty((B, H, c.shape[1]), dtype=torch.float32, device=dev),
                       torch.empty((B, H, c.shape[1]), dtype=torch.float32, device=dev))
    pulls = _post_pulls(be, plan, off, k, v)
    _end_pass(be)
    for idx, st in enumerate(plan.steps):
        bufs, ev = pulls[idx]
        be.wait_event("main", ev)
        for (qi, ki) in st.pairs:
            kb, vb = bufs[ki]
            a = acc[qi] or (None, None, None)
            ops.fwd_step(q_chunks[qi], kb, vb, out_chunks[qi], lse_chunks[qi], a[0], a[1], a[2],
                         plan.q_chunks[qi].pos0, st.kv[ki].pos0, causal, bias, seg,
                         visits[qi][0] == (idx, ki), visits[qi][-1] == (idx, ki))
    out = torch.empty_like(q)
    rx._scatter_q_like(plan, comm, out_chunks, out)
    return out, dict(q_chunks=q_chunks, out_chunks=out_chunks, lse_chunks=lse_chunks)


def run_backward(plan, res, k, v, dout, bias, seg, causal, group, ops, be):
    dev = k.device
    comm = rx._Comm(group, dev)
    q_chunks, out_chunks, lse_chunks = res["q_chunks"], res["out_chunks"], res["lse_chunks"]
    B, Sk, H, D = k.shape
    off, nbytes, max_rows = _layout(plan, k, v)
    be.reserve(nbytes)
    _stage_kv(be, off, k, v)          # the barrier inside also orders this pass after every rank's previous one
    do_chunks = rx._gather_q_like(plan, comm, dout)
    n_q = len(q_chunks)
    delta = [torch.empty_like(l) for l in lse_chunks]
    dq_acc = [torch.zeros(c.shape, dtype=torch.float32, device=dev) for c in q_chunks]
    for qi in range(n_q):
        ops.bwd_prep(out_chunks[qi], do_chunks[qi], delta[qi])
    lse_chunks = [ops.lse_for_bwd(l) for l in lse_chunks]
    dk_acc = torch.zeros(k.shape, dtype=torch.float32, device=dev)
    dv_acc = torch.zeros(v.shape, dtype=torch.float32, device=dev)
    land_shape = (B, max_rows, H, D)
    pulls = _post_pulls(be, plan, off, k, v)
    _end_pass(be)
    slot_tables = {}
    keep = []     # partials stay referenced until the put stream is done with them (they are main-stream allocations)
    for idx, st in enumerate(plan.steps):
        bufs, ev = pulls[idx]
        be.wait_event("main", ev)
        parts = []
        for kv in st.kv:
            if kv.owner == plan.rank:
                parts.append(None)
            else:
                shape = (B, kv.length, H, D)
                parts.append((torch.zeros(shape, dtype=torch.float32, device=dev),
                              torch.zeros(shape, dtype=torch.float32, device=dev)))
        own_views = {}
        for (qi, ki) in st.pairs:
            kv = st.kv[ki]
            kb, vb = bufs[ki]
            if parts[ki] is None:
                if ki not in own_views:
                    if kv.start == 0 and kv.length == Sk:
                        own_views[ki] = (dk_acc, dv_acc, False)
                    else:
                        own_views[ki] = (rx._rows(dk_acc, kv.start, kv.length), rx._rows(dv_acc, kv.start, kv.length), True)
                dkb, dvb = own_views[ki][0], own_views[ki][1]
            else:
                dkb, dvb = parts[ki]
            ops.bwd_step(q_chunks[qi], kb, vb, do_chunks[qi], lse_chunks[qi], delta[qi], dq_acc[qi], dkb, dvb,
                         plan.q_chunks[qi].pos0, kv.pos0, causal, bias, seg)
        for ki, (a, b2, staged) in own_views.items():
            if staged:
                kv = st.kv[ki]
                dk_acc[:, kv.start:kv.start + kv.length].copy_(a)
                dv_acc[:, kv.start:kv.start + kv.length].copy_(b2)
        # one-sided return of this step's partials: put into the owner's landing slot, then signal the owner
        ev = be.record("main")
        be.wait_event("put", ev)
        owners = []
        with be.on_put_stream():
            for ki, kv in enumerate(st.kv):
                if parts[ki] is None:
                    continue
                if kv.owner not in slot_tables:
                    slot_tables[kv.owner] = rs.landing_slots(rs.peer_plan(plan, kv.owner))[0]
                row = slot_tables[kv.owner][(idx, plan.rank, kv.start, kv.length)]
                be.view(kv.owner, off["dk"], land_shape, torch.float32)[:, row:row + kv.length].copy_(parts[ki][0])
                be.view(kv.owner, off["dv"], land_shape, torch.float32)[:, row:row + kv.length].copy_(parts[ki][1])
                if kv.owner not in owners:
                    owners.append(kv.owner)
            for o in owners:
                be.signal(o, 1 + idx)
        keep.append(parts)
    be.wait_event("main", be.record("put"))
    # fold in what the peers returned for my rows, in plan order (mostly landed while I was computing)
    my_slots, _ = rs.landing_slots(plan)
    land_dk = be.view(plan.rank, off["dk"], land_shape, torch.float32)
    land_dv = be.view(plan.rank, off["dv"], land_shape, torch.float32)
    for idx, st in enumerate(plan.steps):
        waited = set()
        for (s, l, peer) in st.sends:
            if peer not in waited:
                be.wait_signal(peer, 1 + idx)
                waited.add(peer)
            row = my_slots[(idx, peer, s, l)]
            ops.accumulate(dk_acc, s, l, land_dk[:, row:row + l])
            ops.accumulate(dv_acc, s, l, land_dv[:, row:row + l])
    dq_chunks = []
    for qi in range(n_q):
        c = torch.empty_like(q_chunks[qi])
        ops.cast(dq_acc[qi], c)
        dq_chunks.append(c)
    dq = torch.empty((B,) + tuple(dout.shape[1:]), dtype=q_chunks[0].dtype, device=dev)
    rx._scatter_q_like(plan, comm, dq_chunks, dq)
    dk = torch.empty_like(k)
    dv = torch.empty_like(v)
    ops.cast(dk_acc, dk)
    ops.cast(dv_acc, dv)
    return dq, dk, dv
