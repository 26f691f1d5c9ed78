"""ONE rank's share of an N-way sequence-parallel pass on ONE GPU, through the real executor (lwm_b200/ring_peer.py)
with a loopback transport: the peers' heaps are local buffers holding random operands, flags are always satisfied, puts
land in scratch. What remains is exactly the rank's own device work — tile kernels, operand staging, carries, partial
folding, local copies — with zero-latency, zero-skew peers; the difference to a real N-GPU pass is exposed communication
and waiting.   python tools/emulate_rank_peer.py [world=8] [S_total=131072] [ranks=0,3,7]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lwm_b200 import ring_peer as rp, ring_schedule as rs, ringattention as ra, synthetic as syn, _lib


class LoopbackTransport(rp.CudaPeerTransport):
    def __init__(self, rank, world, device):
        self.group, self.device, self.rank, self.world = None, device, rank, world
        self.ctx, self.capacity, self.pass_id = None, 0, 0
        self.fan = max(1, int(os.environ.get("LWM_RING_COPY_STREAMS", "4")))
        self.side = {"%s#%d" % (n, i): torch.cuda.Stream(device=device) for n in ("pull", "push") for i in range(self.fan)}
        self._rr = {"pull": 0, "push": 0}
        self.heaps = None

    def ensure(self, nbytes):
        if self.heaps is not None and nbytes <= self.capacity:
            return
        self.capacity = nbytes
        self.own = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        # one shared "peer" heap with random fp16 content (every remote owner reads from it), one scratch for puts
        self.peer = (torch.randn(nbytes // 2, device=self.device, dtype=torch.float16) * 0.7).view(torch.uint8)
        self.scratch = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.heaps = True
        self.pass_id = 0

    def pull(self, dst, peer, off, stream):
        n = dst.numel() * dst.element_size()
        if n == 16:                     # a peer's scale row
            with torch.cuda.stream(self._stream(stream)):
                dst.fill_(2.0 ** -10)
            return
        _lib.call("lwm_ring_copy", _lib.ptr(dst), _lib.ptr(self.peer[off:off + n]), n, self._sp(stream))

    def put(self, src, peer, off, stream):
        n = src.numel() * src.element_size()
        _lib.call("lwm_ring_copy", _lib.ptr(self.scratch[off:off + n]), _lib.ptr(src), n, self._sp(stream))

    def signal(self, peer, flag, value, stream):
        pass

    def wait(self, flag, value, stream):
        pass


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
    ranks = [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["0", str(world // 2 - 1), str(world - 1)])]
    dev = torch.device("cuda", 0)
    Sl, H, D = S // world, 32, 128
    ops = ra.PeerOpsF16
    for rank in ranks:
        tr = LoopbackTransport(rank, world, dev)
        q, k, v, do = [syn.shard(n_, rank, Sl, H, D).to(dev) for n_ in ("q", "k", "v", "do")]
        plan = rs.make_peer_plan(world, rank, Sl, Sl, True, "zigzag")

        def step():
            out, res = rp.run_forward(plan, q, k, v, None, None, True, ops, tr, False)
            return rp.run_backward(plan, res, k, v, do, None, None, True, ops, tr, False)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        n = 5
        for _ in range(n):
            step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        tr.trace = []
        t0 = torch.cuda.Event(enable_timing=True)
        t0.record()
        step()
        torch.cuda.synchronize()
        spans = tr.span_times(t0, tr.trace)
        tr.trace = None
        kern = sum(e - s for (lab, st, s, e) in spans if "kernel" in lab)
        fwd_k = sum(e - s for (lab, st, s, e) in spans if "fwd kernel" in lab)
        print("rank %d of %d (S=%d): %.2f ms per fwd+bwd pass; tile kernels %.2f ms (fwd %.2f, bwd %.2f)" % (
            rank, world, S, ms, kern, fwd_k, kern - fwd_k), flush=True)
        for (lab, st, s, e) in sorted(spans, key=lambda x: x[2]):
            if "kernel" not in lab or rank == ranks[0]:
                print("    %-5s %8.3f .. %8.3f (%7.3f)  %s" % (st, s, e, e - s, lab))
        del tr, q, k, v, do
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
