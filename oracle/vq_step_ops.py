"""CPU stand-ins for the VQGAN C-ABI calls — TEST INFRASTRUCTURE ONLY (like oracle/step_ops.py for the ring).

Same method surface as lwm_b200.vqgan.Ops, implemented with torch-CPU float32 primitives on the SAME packed operands the
CUDA kernels consume ([N,H,W,Cpad] operand planes, PackedConv weights [tap][Cout_pad][Cpad] as hi + lo bf16), so that
the product's host logic — flax parameter-tree traversal and auto-naming, weight packing and channel padding, residual /
shortcut wiring, the (0,1) Downsample pad, nearest 2x resize, the video reshape — runs on CPU (tests/test_vqgan_host_cpu.py)
against the fixture produced by executing the reference module."""
import torch
import torch.nn.functional as F

from . import vqgan_ref as vr


class CpuVqOps:
    n_pass = 3

    def gn_stats(self, x):
        raise NotImplementedError("folded into prep on CPU")

    def prep(self, x, gn=None, upsample=False, cpad=None):
        N, H, W, C = x.shape
        cpad = cpad or -(-C // 64) * 64
        y = x.float()
        if gn is not None:
            y = vr.silu(vr.group_norm(y, gn))
        if upsample:
            y = y.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
        return F.pad(y, (0, cpad - C)), None          # zero channel padding, like the kernel's planes

    def conv_gn(self, x, pc, gn=None, upsample=False, stride=1, residual=None, clip=False, want_stats=False):
        return self.conv(self.prep(x, gn, upsample), pc, stride=stride, residual=residual, clip=clip)

    def conv(self, planes, pc, stride=1, residual=None, clip=False, want_stats=False):
        a, _ = planes
        assert a.shape[-1] == pc.cpad, (a.shape, pc.cpad)
        w = (pc.w_hi.float() + pc.w_lo.float()).reshape(pc.k, pc.k, pc.cout_pad, pc.cpad)    # [ky][kx][Cout_pad][Cpad]
        wc = w.permute(2, 3, 0, 1)                                                          # OIHW
        x = a.permute(0, 3, 1, 2)
        if stride == 1:
            y = F.conv2d(x, wc, padding=pc.k // 2)
        else:                                   # Downsample: implicit zero row/column at the bottom/right, VALID
            y = F.conv2d(F.pad(x, (0, 1, 0, 1)), wc, stride=2)
        y = y.permute(0, 2, 3, 1)[..., :pc.cout] + pc.bias
        if residual is not None:
            y = y + residual
        return y.clamp(-1.0, 1.0) if clip else y

    def conv_cin3(self, x, pc):
        return vr.conv2d(x.float(), {"kernel": pc.w_hwio, "bias": pc.bias})

    def vq_argmin(self, z_flat, emb, want_zq=True):
        zq, idx = vr.vector_quantize(z_flat.numpy(), emb.numpy())
        return torch.from_numpy(zq), torch.from_numpy(idx)

    def vq_gather(self, idx_flat, emb):
        return emb[idx_flat.long()]
