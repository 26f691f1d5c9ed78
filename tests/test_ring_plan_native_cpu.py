"""The schedule and the heap layout of the peer-memory ring exist twice: in Python (lwm_b200/ring_schedule.py::
make_peer_plan, lwm_b200/ring_peer.py::Layout — what the executor runs on) and behind the C ABI (lwm_ring_plan /
lwm_ring_layout, for hosts in other languages). They are pure functions: this test compares them field by field."""
import ctypes

import pytest

LL, I = ctypes.c_longlong, ctypes.c_int
MAXC, MAXL = 32, 128


class Chunk(ctypes.Structure):
    _fields_ = [("owner", I), ("index", I), ("start", LL), ("length", LL), ("pos0", LL)]


class Launch(ctypes.Structure):
    _fields_ = [("q_chunk", I), ("owner", I), ("key_row0", LL), ("rows", LL)]


class QSend(ctypes.Structure):
    _fields_ = [("start", LL), ("length", LL), ("peer", I)]


class Incoming(ctypes.Structure):
    _fields_ = [("chunk_index", I), ("peer", I)]


class Plan(ctypes.Structure):
    _fields_ = [("world", I), ("rank", I), ("zigzag", I), ("chunks_per_rank", I),
                ("n_q", I), ("q", Chunk * 2),
                ("n_q_sends", I), ("q_sends", QSend * MAXC),
                ("n_fwd_groups", I), ("n_fwd_chunks", I), ("n_fwd_launches", I),
                ("fwd_group_first_chunk", I * (MAXC + 1)), ("fwd_group_first_launch", I * (MAXC + 1)),
                ("fwd_chunks", Chunk * MAXC), ("fwd_launches", Launch * MAXL),
                ("n_bwd_groups", I), ("n_bwd_chunks", I), ("n_bwd_launches", I),
                ("bwd_group_first_chunk", I * (MAXC + 1)), ("bwd_group_first_launch", I * (MAXC + 1)),
                ("bwd_chunks", Chunk * MAXC), ("bwd_launches", Launch * MAXL),
                ("n_incoming", I), ("incoming", Incoming * MAXC),
                ("n_own", I), ("own_computed", I * 2)]


class Layout(ctypes.Structure):
    _fields_ = [(n, LL) for n in ("scales", "kg", "vg", "qs", "lq4", "lq2", "lp", "slot_bytes", "set_bytes", "total",
                                  "chunk_rows")] + [("n_slots", I)]


def _groups(plan, which):
    n = getattr(plan, "n_%s_groups" % which)
    fc, fl = getattr(plan, "%s_group_first_chunk" % which), getattr(plan, "%s_group_first_launch" % which)
    chunks, launches = getattr(plan, "%s_chunks" % which), getattr(plan, "%s_launches" % which)
    out = []
    for g in range(n):
        cs = [(c.owner, c.index, c.start, c.length, c.pos0) for c in chunks[fc[g]:fc[g + 1]]]
        ls = [(l.q_chunk, l.key_row0, l.rows, l.owner) for l in launches[fl[g]:fl[g + 1]]]
        out.append((cs, ls))
    return out


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8, 16])
@pytest.mark.parametrize("layout,causal", [("zigzag", True), ("contiguous", True), ("contiguous", False)])
def test_native_plan_equals_python_plan(lib, world, layout, causal):
    from lwm_b200 import ring_schedule as rs
    S = 1024
    for rank in range(world):
        for fgc in (2, 4):
            p = Plan()
            assert lib.lwm_ring_plan(world, rank, S, S, int(causal), int(layout == "zigzag"), fgc, ctypes.byref(p)) == 0, \
                lib.lwm_last_error()
            ref = rs.make_peer_plan(world, rank, S, S, causal, layout, fwd_group_chunks=fgc)
            assert (p.world, p.rank, p.chunks_per_rank) == (ref.world, ref.rank, ref.chunks_per_rank)
            assert [(c.owner, c.start, c.length, c.pos0) for c in p.q[:p.n_q]] == \
                [(q.owner, q.start, q.length, q.pos0) for q in ref.q_chunks]
            assert [(s.start, s.length, s.peer) for s in p.q_sends[:p.n_q_sends]] == list(ref.q_sends)
            for which, groups in (("fwd", ref.fwd_groups), ("bwd", ref.bwd_groups)):
                want = [([(c.owner, c.index, c.start, c.length, c.pos0) for c in g.chunks], list(g.launches)) for g in groups]
                assert _groups(p, which) == want, (world, rank, which)
            assert [(x.chunk_index, x.peer) for x in p.incoming[:p.n_incoming]] == list(ref.incoming)
            assert list(p.own_computed[:p.n_own]) == list(ref.own_computed)


def test_native_plan_rejects_bad_arguments(lib):
    p = Plan()
    assert lib.lwm_ring_plan(8, 8, 1024, 1024, 1, 1, 4, ctypes.byref(p)) == 3
    assert lib.lwm_ring_plan(8, 0, 1000, 1000, 1, 1, 4, ctypes.byref(p)) == 2 and b"zigzag" in lib.lwm_last_error()
    assert lib.lwm_ring_plan(8, 0, 1024, 1024, 1, 1, 4, None) == 3


@pytest.mark.parametrize("B,Sq,Sk,world,cpr,isz", [(1, 16384, 16384, 8, 2, 2), (2, 512, 512, 4, 2, 4), (1, 256, 1024, 3, 1, 2)])
def test_native_layout_equals_python_layout(lib, B, Sq, Sk, world, cpr, isz):
    from lwm_b200.ring_peer import Layout as PyLayout
    lay = Layout()
    assert lib.lwm_ring_layout(B, Sq, Sk, 32, 128, world, cpr, isz, ctypes.byref(lay)) == 0
    ref = PyLayout(B, Sq, Sk, 32, 128, world, cpr, isz)
    assert (lay.scales, lay.kg, lay.vg, lay.qs, lay.lq4, lay.lq2, lay.lp) == (ref.abs, ref.kg, ref.vg, ref.qs, ref.lq4,
                                                                             ref.lq2, ref.lp)
    assert (lay.slot_bytes, lay.set_bytes, lay.total, lay.chunk_rows, lay.n_slots) == (ref.slot_bytes, ref.set_bytes,
                                                                                       ref.total, ref.chunk_rows, ref.n_slots)
