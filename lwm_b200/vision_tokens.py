"""Vision token framing around the VQGAN — host mirror of the wire format in lwm/vision_chat.py:87-105,
lwm/data.py:193-219 and lwm/vision_generation.py:160,221 over the CUDA kernels lwm_vq_frame_tokens /
lwm_vq_unframe_tokens: codebook indices stay on the device from VQGAN.encode to the LM token buffer."""
import numpy as np
import torch

from . import _lib

EOF_TOKEN = 8192             # end of each frame (data.py:134)
EOV_TOKEN = 8193             # end of vision generation (data.py:135)
N_TOKENS_PER_FRAME = 256     # 16 x 16 VQ codes (data.py:136)


def select_frames(n_frames, max_n_frames):
    """data.py:196-202: uniform frame selection, or None when every frame is kept."""
    if max_n_frames > 0 and n_frames > max_n_frames:
        return np.linspace(0, n_frames - 1, max_n_frames).astype(int)
    return None


def frame_tokens(codes, max_n_frames=-1, eof_token=EOF_TOKEN, eov_token=EOV_TOKEN, eov_every_frame=False):
    """codes int32 [T,16,16] / [T,P] (one clip) or [B,T,16,16] / [B,T,P] -> tokens int32 [B?, T'*(P+1)]: every frame's
    codes then eof_token, eov_token after the clip's last frame — the training data format (data.py:206-212).
    eov_every_frame=True reproduces the stream lwm/vision_chat.py:91-107 actually feeds the model: it encodes ONE frame
    per call (B = 1), so its `t == len(enc) - 1` test is true for every frame and 8193 follows each of them."""
    if eov_every_frame:
        eof_token = eov_token
    if not codes.is_cuda:
        raise _lib.LwmError("frame_tokens: codes must be a CUDA tensor (no CPU path)")
    c = codes.to(torch.int32)
    if c.dim() == 4:                       # [B,T,h,w]
        c = c.reshape(c.shape[0], c.shape[1], -1)
        batched = True
    elif c.dim() == 3 and codes.shape[-1] == codes.shape[-2] and codes.shape[-1] * codes.shape[-2] == N_TOKENS_PER_FRAME:
        c = c.reshape(1, c.shape[0], -1)   # [T,16,16]
        batched = False
    elif c.dim() == 3:                     # [B,T,P]
        batched = True
    elif c.dim() == 2:                     # [T,P]
        c = c.unsqueeze(0)
        batched = False
    else:
        raise _lib.LwmError("frame_tokens: codes must be [T,h,w], [T,P], [B,T,h,w] or [B,T,P]")
    c = c.contiguous()
    B, T, P = c.shape
    if T == 0:
        raise _lib.LwmError("frame_tokens: a clip needs at least one frame")    # data.py:205
    sel = select_frames(T, max_n_frames)
    T_out = T if sel is None else len(sel)
    idx = None if sel is None else torch.from_numpy(sel.astype(np.int32)).to(c.device)
    out = torch.empty(B, T_out * (P + 1), dtype=torch.int32, device=c.device)
    _lib.call("lwm_vq_frame_tokens", _lib.ptr(c), _lib.ptr(idx), _lib.ptr(out), B, T, T_out, P, int(eof_token),
              int(eov_token), _lib.stream_ptr())
    return out if batched else out[0]


def _as_frames(t, P):
    """view a token tensor as [..., n_frames, P+1]; a flat (1-D) stream is always split into frames"""
    if t.dim() == 0 or t.shape[-1] % (P + 1):
        raise _lib.LwmError("unframe_tokens: last dimension must be a multiple of tokens_per_frame + 1")
    if t.dim() == 1 or t.shape[-1] != P + 1:
        t = t.reshape(*t.shape[:-1], t.shape[-1] // (P + 1), P + 1)
    return t


def unframe_tokens(tokens, n_tokens_per_frame=N_TOKENS_PER_FRAME, grid=(16, 16)):
    """tokens int [..., T*(P+1)] or [..., T, P+1] -> codes int32 [..., T, 16, 16] with the per-frame delimiter dropped
    (vision_generation.py:160 `output[:, :-1].reshape(-1,16,16)`, :221 `output[:, :, :-1].reshape(-1,n_frames,16,16)`)."""
    if not tokens.is_cuda:
        raise _lib.LwmError("unframe_tokens: tokens must be a CUDA tensor (no CPU path)")
    P = n_tokens_per_frame
    t = _as_frames(tokens.to(torch.int32), P).contiguous()
    lead = t.shape[:-1]
    n = int(np.prod(lead)) if len(lead) else 1
    out = torch.empty(*lead, P, dtype=torch.int32, device=t.device)
    _lib.call("lwm_vq_unframe_tokens", _lib.ptr(t), _lib.ptr(out), n, P, _lib.stream_ptr())
    if grid is not None and grid[0] * grid[1] == P:
        out = out.reshape(*lead, *grid)
    return out


def vision_mask(n_frames, n_start, n_end, n_tokens_per_frame=N_TOKENS_PER_FRAME):
    """vision_mask of one `vision` field (data.py:216-219): False over <vision>, True over codes AND delimiters,
    False over </vision>."""
    return [False] * n_start + [True] * (n_tokens_per_frame * n_frames + n_frames) + [False] * n_end
