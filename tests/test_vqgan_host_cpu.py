"""CPU test of the VQGAN product's HOST logic (lwm_b200/vqgan.py) with the CUDA calls replaced by CPU stand-ins that
consume the same packed operands (oracle/vq_step_ops.py): parameter-tree traversal with flax auto-names (incl. the
reversed UpsamplingBlock numbering), PackedConv weight layout and channel padding, shortcut / residual wiring, Downsample
and Upsample placement, quant / post-quant convs, the [B,T,...] video branch — checked against the fixture produced by
EXECUTING the reference module (tests/golden/vqgan_reference_small.npz)."""
import os

import numpy as np
import torch

from helpers import rel_fro

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _model():
    from lwm_b200.vqgan import VQGANConfig, VQGANModel
    from oracle import vqgan_ref as vr
    from oracle.vq_step_ops import CpuVqOps
    gold = np.load(os.path.join(GOLD, "vqgan_reference_small.npz"))
    cfgd = dict(resolution=int(gold["cfg_resolution"]), hidden_channels=int(gold["cfg_hidden"]),
                num_embeddings=int(gold["cfg_codes"]))
    params = vr.init_params(cfgd, seed=int(gold["param_seed"]), codebook="normal")
    model = VQGANModel(VQGANConfig.get_default_config(cfgd), params, device="cpu")
    model.ops = CpuVqOps()
    model._to_dev = lambda x: torch.as_tensor(np.asarray(x)).float()      # the product refuses to run without a GPU
    return model, gold


def test_host_logic_encode_matches_reference_fixture():
    model, gold = _model()
    zq, idx = model.encode(gold["pixels"].reshape(1, 2, 64, 64, 3))       # video branch (vqgan.py:118-127)
    assert tuple(idx.shape) == gold["idx"].shape and tuple(zq.shape) == gold["zq"].shape
    # weights go through the hi+lo bf16 split (2^-16 relative), so a near-tied code may differ from the fp32 reference
    assert (idx.numpy() == gold["idx"]).mean() >= 0.99
    same = idx.numpy() == gold["idx"]
    assert rel_fro(zq.numpy()[same], gold["zq"][same]) < 1e-4


def test_host_logic_decode_matches_reference_fixture():
    model, gold = _model()
    rec = model.decode(gold["codes"])
    assert tuple(rec.shape) == gold["recon"].shape
    assert rel_fro(rec.numpy(), gold["recon"]) < 1e-4
    assert float(rec.max()) <= 1.0 and float(rec.min()) >= -1.0


def test_packed_conv_layout_and_padding():
    """[tap][Cout_pad][Cpad], tap = ky*3+kx, zero padded; hi + lo reproduces the fp32 kernel to 2^-16"""
    from lwm_b200.vqgan import PackedConv
    g = torch.Generator().manual_seed(0)
    w = torch.randn(3, 3, 96, 40, generator=g)
    pc = PackedConv({"kernel": w, "bias": torch.zeros(40)}, torch.device("cpu"))
    assert (pc.k, pc.cin, pc.cout, pc.cpad, pc.cout_pad) == (3, 96, 40, 128, 48)
    full = pc.w_hi.float() + pc.w_lo.float()
    assert full.shape == (9, 48, 128)
    assert torch.all(full[:, 40:, :] == 0) and torch.all(full[:, :, 96:] == 0)
    for ky in range(3):
        for kx in range(3):
            assert torch.allclose(full[ky * 3 + kx, :40, :96], w[ky, kx].T, rtol=0, atol=2.0 ** -15 * float(w.abs().max()))


def test_packed_conv_stacked_fp16_weights():
    """fp16x2 scheme: [tap][Cout_pad/BN][hi rows | lo rows][Cpad], (hi + lo) * w_scale_inv reproduces the fp32 kernel"""
    from lwm_b200.vqgan import PackedConv
    g = torch.Generator().manual_seed(1)
    for cin, cout in ((96, 40), (128, 256), (64, 768)):
        w = torch.randn(3, 3, cin, cout, generator=g) * 0.03
        pc = PackedConv({"kernel": w, "bias": torch.zeros(cout)}, torch.device("cpu"))
        nt = pc.cout_pad // pc.bn
        assert pc.cout_pad % pc.bn == 0 and pc.bn % 16 == 0 and pc.bn <= 128
        assert tuple(pc.w_stack.shape) == (9, nt, 2 * pc.bn, pc.cpad) and pc.w_stack.dtype == torch.float16
        hi, lo = pc.w_stack[:, :, :pc.bn].float(), pc.w_stack[:, :, pc.bn:].float()
        assert float(hi.abs().max()) < 8192.0 * 1.01 and float(hi.abs().max()) >= 4096.0     # |w|max lands in [2^12, 2^13)
        full = ((hi + lo) * pc.w_scale_inv).reshape(9, pc.cout_pad, pc.cpad)
        for ky in range(3):
            for kx in range(3):
                assert torch.allclose(full[ky * 3 + kx, :cout, :cin], w[ky, kx].T, rtol=0,
                                      atol=2.0 ** -20 * float(w.abs().max()))
        assert torch.all(full[:, cout:, :] == 0) and torch.all(full[:, :, cin:] == 0)


def _replicate_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lwm_b200.vqgan import VQGAN
        from oracle import vqgan_ref as vr
        from oracle.vq_step_ops import CpuVqOps
        gold = np.load(os.path.join(GOLD, "vqgan_reference_small.npz"))
        cfgd = dict(resolution=int(gold["cfg_resolution"]), hidden_channels=int(gold["cfg_hidden"]),
                    num_embeddings=int(gold["cfg_codes"]))
        params = vr.init_params(cfgd, seed=int(gold["param_seed"]), codebook="normal")
        from lwm_b200.vqgan import VQGANConfig, VQGANModel
        tok = VQGAN.__new__(VQGAN)
        tok.replicate, tok.group = True, None
        tok.model = VQGANModel(VQGANConfig.get_default_config(cfgd), params, device="cpu")
        tok.model.ops = CpuVqOps()
        tok.model._to_dev = lambda x: torch.as_tensor(np.asarray(x)).float()
        px = torch.as_tensor(gold["pixels"]).reshape(2, 1, 64, 64, 3)         # leading axis = 2 ranks, one frame each
        zq, idx = tok.encode(px)
        ret[rank] = (tuple(idx.shape), idx.numpy().reshape(-1).tolist())
        try:
            tok.encode(px[:1])
            ret["raised_%d" % rank] = False
        except ValueError:
            ret["raised_%d" % rank] = True
    finally:
        dist.destroy_process_group()


def test_replicate_maps_the_leading_axis_over_ranks():
    """VQGAN(replicate=True): the reference's jax.pmap branch (lwm/vqgan.py:20-28) over a 2-rank gloo group"""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ret = mp.Manager().dict()
    mp.spawn(_replicate_worker, args=(2, port, ret), nprocs=2, join=True)
    gold = np.load(os.path.join(GOLD, "vqgan_reference_small.npz"))
    for r in range(2):
        shape, flat = ret[r]
        assert shape == (2, 1) + tuple(gold["idx"].shape[-2:])
        assert (np.asarray(flat).reshape(gold["idx"].shape) == gold["idx"]).mean() >= 0.99
        assert ret["raised_%d" % r]
    assert ret[0][1] == ret[1][1]
