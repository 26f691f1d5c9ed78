"""CPU restatement of the reference VQGAN tokenizer (lwm/vqgan.py:105-351) with the flax / jax
default semantics it relies on: NHWC activations, HWIO conv kernels with
'SAME' padding, GroupNorm(32 groups, eps 1e-6, fast variance E[x^2]-E[x]^2 clamped at 0),
silu = x*sigmoid(x), Downsample = zero-pad bottom/right by one then 3x3 stride-2 VALID conv
(vqgan.py:292-300), Upsample = nearest 2x (out[i,j] = in[i//2,j//2]) then 3x3 conv
(vqgan.py:312-318), VectorQuantizer distance d = sum z^2 + sum e^2 - 2 z.e with first-index argmin
(vqgan.py:207-212), straight-through z + (z_q - z) (vqgan.py:215), decode clip to [-1,1]
(vqgan.py:141). Parameter tree uses flax's auto-naming (Conv_0, GroupNorm_0, ResnetBlock_0, ...).

Convolutions / GroupNorm run through torch CPU float32 (or float64 with dtype=torch.float64 to get
a higher-precision yardstick); the VectorQuantizer is evaluated in numpy float32 with a FIXED
operation order (sequential over the 64 dims, separate multiply and add roundings) which the CUDA
kernel replicates bit for bit — see `vq_distances_f32`.

TEST INFRASTRUCTURE ONLY. PARITY UNPINNED against the reference binary (jax/flax are not
installable offline and the reference ships no golden vectors) — see oracle/__init__.py.
"""
import numpy as np
import torch
import torch.nn.functional as F

DEFAULT_CONFIG = dict(  # VQGANConfig defaults, lwm/vqgan.py:62-77
    resolution=256, num_channels=3, hidden_channels=128, channel_mult=(1, 2, 2, 4, 6), num_res_blocks=2,
    attn_resolutions=(), no_attn_mid_block=True, z_channels=64, num_embeddings=8192, quantized_embed_dim=64,
    dropout=0.0, resample_with_conv=True, commitment_cost=0.25)


# ------------------------------------------------------------------------------------------------
# synthetic parameters in the flax tree layout
# ------------------------------------------------------------------------------------------------
def _conv_p(g, k, cin, cout):
    w = torch.randn(k, k, cin, cout, generator=g) * (1.0 / np.sqrt(k * k * cin))
    b = torch.randn(cout, generator=g) * 0.02
    return {"kernel": w, "bias": b}


def _gn_p(g, c):
    return {"scale": 1.0 + 0.02 * torch.randn(c, generator=g), "bias": 0.02 * torch.randn(c, generator=g)}


def _resnet_p(g, cin, cout):
    p = {"GroupNorm_0": _gn_p(g, cin), "Conv_0": _conv_p(g, 3, cin, cout), "GroupNorm_1": _gn_p(g, cout),
         "Conv_1": _conv_p(g, 3, cout, cout)}
    if cin != cout:
        p["Conv_2"] = _conv_p(g, 1, cin, cout)  # nin shortcut (use_conv_shortcut=False, vqgan.py:258-262)
    return p


def init_params(config=None, seed=0, codebook="normal"):
    """Random parameters with the reference's shapes and flax names. codebook: 'uniform' is the
    reference initialiser U(-1/n_e, 1/n_e) (vqgan.py:198-200); 'normal' a realistic N(0,1) spread."""
    cfg = dict(DEFAULT_CONFIG, **(config or {}))
    g = torch.Generator().manual_seed(seed)
    hc, mult, nres = cfg["hidden_channels"], cfg["channel_mult"], cfg["num_res_blocks"]
    nlev = len(mult)
    enc = {"Conv_0": _conv_p(g, 3, cfg["num_channels"], hc)}
    cin = hc
    for i in range(nlev):
        cout = hc * mult[i]
        blk = {}
        for j in range(nres):
            blk["ResnetBlock_%d" % j] = _resnet_p(g, cin, cout)
            cin = cout
        if i != nlev - 1:
            blk["Downsample_0"] = {"Conv_0": _conv_p(g, 3, cin, cin)}
        enc["DownsamplingBlock_%d" % i] = blk
    enc["MidBlock_0"] = {"ResnetBlock_0": _resnet_p(g, cin, cin), "ResnetBlock_1": _resnet_p(g, cin, cin)}
    enc["GroupNorm_0"] = _gn_p(g, cin)
    enc["Conv_1"] = _conv_p(g, 3, cin, cfg["z_channels"])

    ctop = hc * mult[-1]
    dec = {"Conv_0": _conv_p(g, 3, cfg["z_channels"], ctop),
           "MidBlock_0": {"ResnetBlock_0": _resnet_p(g, ctop, ctop), "ResnetBlock_1": _resnet_p(g, ctop, ctop)}}
    cin = ctop
    for n, i in enumerate(reversed(range(nlev))):   # UpsamplingBlock_0 <=> block_idx = nlev-1 (vqgan.py:179-180)
        cout = hc * mult[i]
        blk = {}
        for j in range(nres + 1):
            blk["ResnetBlock_%d" % j] = _resnet_p(g, cin, cout)
            cin = cout
        if i != 0:
            blk["Upsample_0"] = {"Conv_0": _conv_p(g, 3, cin, cin)}
        dec["UpsamplingBlock_%d" % n] = blk
    dec["GroupNorm_0"] = _gn_p(g, cin)
    dec["Conv_1"] = _conv_p(g, 3, cin, cfg["num_channels"])

    n_e, e_dim = cfg["num_embeddings"], cfg["quantized_embed_dim"]
    if codebook == "uniform":
        emb = (torch.rand(n_e, e_dim, generator=g) * 2 - 1) / n_e
    else:
        emb = torch.randn(n_e, e_dim, generator=g)
    return {"encoder": enc, "decoder": dec, "quantize": {"embeddings": emb},
            "quant_conv": _conv_p(g, 1, cfg["z_channels"], e_dim),
            "post_quant_conv": _conv_p(g, 1, e_dim, cfg["z_channels"])}


# ------------------------------------------------------------------------------------------------
# layers (NHWC in / NHWC out)
# ------------------------------------------------------------------------------------------------
def conv2d(x, p, stride=1, padding="SAME", round_fn=None):
    """flax nn.Conv: x [N,H,W,Cin], kernel HWIO. round_fn (optional) emulates a reduced-precision
    tensor-core operand format on both inputs (used to size the CUDA path's precision modes)."""
    w = p["kernel"].to(x.dtype)
    k = w.shape[0]
    xin, win = x, w
    if round_fn is not None:
        xin, win = round_fn(x), round_fn(w)
    xc = xin.permute(0, 3, 1, 2)
    wc = win.permute(3, 2, 0, 1)  # OIHW
    pad = (k // 2) if padding == "SAME" else 0
    y = F.conv2d(xc, wc, bias=None, stride=stride, padding=pad)
    return y.permute(0, 2, 3, 1) + p["bias"].to(x.dtype)


def group_norm(x, p, groups=32, eps=1e-6):
    """flax nn.GroupNorm(): statistics over (H, W, C/groups); var = max(0, E[x^2] - E[x]^2)."""
    N, H, W, C = x.shape
    xg = x.reshape(N, H * W, groups, C // groups)
    mean = xg.mean(dim=(1, 3), keepdim=True)
    msq = (xg * xg).mean(dim=(1, 3), keepdim=True)
    var = torch.clamp(msq - mean * mean, min=0.0)
    y = (xg - mean) * torch.rsqrt(var + eps)
    return y.reshape(N, H, W, C) * p["scale"].to(x.dtype) + p["bias"].to(x.dtype)


def silu(x):
    return x * torch.sigmoid(x)


def resnet_block(x, p, round_fn=None):
    """ResnetBlock, vqgan.py:242-263 (dropout is p=0 / deterministic)."""
    h = conv2d(silu(group_norm(x, p["GroupNorm_0"])), p["Conv_0"], round_fn=round_fn)
    h = conv2d(silu(group_norm(h, p["GroupNorm_1"])), p["Conv_1"], round_fn=round_fn)
    res = conv2d(x, p["Conv_2"], round_fn=round_fn) if "Conv_2" in p else x
    return h + res


def downsample(x, p, round_fn=None):
    """vqgan.py:292-300: pad bottom/right by one, 3x3 stride-2 VALID conv."""
    x = F.pad(x, (0, 0, 0, 1, 0, 1))  # NHWC: (C: 0,0) (W: 0,1) (H: 0,1)
    return conv2d(x, p["Conv_0"], stride=2, padding="VALID", round_fn=round_fn)


def upsample(x, p, round_fn=None):
    """vqgan.py:310-318: nearest 2x (out[i,j] = in[i//2, j//2]) then 3x3 SAME conv."""
    x = x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    return conv2d(x, p["Conv_0"], round_fn=round_fn)


def encoder(x, p, cfg=None, round_fn=None):
    """Encoder.__call__, vqgan.py:153-164."""
    cfg = dict(DEFAULT_CONFIG, **(cfg or {}))
    nlev = len(cfg["channel_mult"])
    h = conv2d(x, p["Conv_0"], round_fn=round_fn)
    for i in range(nlev):
        blk = p["DownsamplingBlock_%d" % i]
        for j in range(cfg["num_res_blocks"]):
            h = resnet_block(h, blk["ResnetBlock_%d" % j], round_fn)
        if i != nlev - 1:
            h = downsample(h, blk["Downsample_0"], round_fn)
    h = resnet_block(h, p["MidBlock_0"]["ResnetBlock_0"], round_fn)
    h = resnet_block(h, p["MidBlock_0"]["ResnetBlock_1"], round_fn)
    h = silu(group_norm(h, p["GroupNorm_0"]))
    return conv2d(h, p["Conv_1"], round_fn=round_fn)


def decoder(z, p, cfg=None, round_fn=None):
    """Decoder.__call__, vqgan.py:171-184."""
    cfg = dict(DEFAULT_CONFIG, **(cfg or {}))
    nlev = len(cfg["channel_mult"])
    h = conv2d(z, p["Conv_0"], round_fn=round_fn)
    h = resnet_block(h, p["MidBlock_0"]["ResnetBlock_0"], round_fn)
    h = resnet_block(h, p["MidBlock_0"]["ResnetBlock_1"], round_fn)
    for n, i in enumerate(reversed(range(nlev))):
        blk = p["UpsamplingBlock_%d" % n]
        for j in range(cfg["num_res_blocks"] + 1):
            h = resnet_block(h, blk["ResnetBlock_%d" % j], round_fn)
        if i != 0:
            h = upsample(h, blk["Upsample_0"], round_fn)
    h = silu(group_norm(h, p["GroupNorm_0"]))
    return conv2d(h, p["Conv_1"], round_fn=round_fn)


# ------------------------------------------------------------------------------------------------
# VectorQuantizer (vqgan.py:187-221) with a pinned float32 operation order
# ------------------------------------------------------------------------------------------------
def _seq_sum_sq_f32(a):
    """sum_d a[:, d]^2 in float32, d ascending, product and sum rounded separately."""
    acc = np.zeros(a.shape[0], dtype=np.float32)
    for d in range(a.shape[1]):
        acc = (acc + (a[:, d] * a[:, d]).astype(np.float32)).astype(np.float32)
    return acc


def vq_distances_f32(z_flat, emb):
    """d[b,n] = (sum_d z^2 + sum_d e^2) - 2 * sum_d z e   (association order of vqgan.py:208-210),
    every operation rounded to float32, dot product accumulated sequentially over d = 0..63 with a
    separately rounded multiply and add (no FMA). The CUDA kernel uses exactly this order."""
    z = np.ascontiguousarray(z_flat, dtype=np.float32)
    e = np.ascontiguousarray(emb, dtype=np.float32)
    zz = _seq_sum_sq_f32(z)
    ee = _seq_sum_sq_f32(e)
    dot = np.zeros((z.shape[0], e.shape[0]), dtype=np.float32)
    for d in range(z.shape[1]):
        dot = (dot + (z[:, d:d + 1] * e[None, :, d]).astype(np.float32)).astype(np.float32)
    return ((zz[:, None] + ee[None, :]).astype(np.float32) - (np.float32(2.0) * dot)).astype(np.float32)


def vector_quantize(z, emb):
    """z [..., e_dim] float32 -> (z_q straight-through value, indices int32)."""
    z = np.asarray(z, dtype=np.float32)
    emb = np.asarray(emb, dtype=np.float32)
    flat = z.reshape(-1, z.shape[-1])
    idx = np.empty(flat.shape[0], dtype=np.int32)
    for s in range(0, flat.shape[0], 512):
        idx[s:s + 512] = np.argmin(vq_distances_f32(flat[s:s + 512], emb), axis=1).astype(np.int32)
    zq = emb[idx].reshape(z.shape)
    st = (z + (zq - z).astype(np.float32)).astype(np.float32)   # z + stop_gradient(z_q - z)
    return st, idx.reshape(z.shape[:-1])


def encode(pixel_values, params, cfg=None, round_fn=None):
    """VQGANModel.encode, vqgan.py:117-128. pixel_values [N,H,W,3] or [B,T,H,W,3] in [-1,1]."""
    x = torch.as_tensor(pixel_values, dtype=torch.float32)
    T = None
    if x.dim() == 5:
        T = x.shape[1]
        x = x.reshape((-1,) + tuple(x.shape[2:]))
    h = encoder(x, params["encoder"], cfg, round_fn)
    h = conv2d(h, params["quant_conv"], round_fn=round_fn)
    zq, idx = vector_quantize(h.numpy(), params["quantize"]["embeddings"].numpy())
    if T is not None:
        zq = zq.reshape((-1, T) + zq.shape[1:])
        idx = idx.reshape((-1, T) + idx.shape[1:])
    return zq, idx, h.numpy()


def decode(indices, params, cfg=None, round_fn=None):
    """VQGANModel.decode, vqgan.py:130-141."""
    idx = np.asarray(indices)
    emb = params["quantize"]["embeddings"]
    z = emb[torch.as_tensor(idx, dtype=torch.long)]
    T = None
    if z.dim() == 5:
        T = z.shape[1]
        z = z.reshape((-1,) + tuple(z.shape[2:]))
    h = conv2d(z, params["post_quant_conv"], round_fn=round_fn)
    y = decoder(h, params["decoder"], cfg, round_fn)
    if T is not None:
        y = y.reshape((-1, T) + tuple(y.shape[1:]))
    return torch.clamp(y, -1.0, 1.0).numpy()


def round_bf16(x):
    return x.to(torch.bfloat16).to(x.dtype)


def round_bf16x2(x):
    """hi + lo split into two bf16 values (what the 'bf16x3' tensor-core mode feeds the MMA)."""
    hi = x.to(torch.bfloat16).to(x.dtype)
    lo = (x - hi).to(torch.bfloat16).to(x.dtype)
    return hi + lo
