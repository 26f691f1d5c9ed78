"""CPU oracle of the vision token framing — TEST INFRASTRUCTURE ONLY (never imported by the product).
Plain-Python restatement of lwm/vision_chat.py:97-104 (encode loop tail) and lwm/data.py:193-219 (the `vision` field
branch of VisionTextProcessor.__call__), and of the un-framing at lwm/vision_generation.py:160,221.

Parity status: PINNED — tests/golden/vision_tokens_reference.npz holds the output of the reference's own
VisionTextProcessor class executed here (tools/make_golden_next_rows_from_reference.py)."""
import numpy as np


def frame_tokens(vision_tokens, n_tokens_per_frame=256, eof_token=8192, eov_token=8193, max_n_frames=-1):
    """data.py:193-212 without the <vision> / </vision> text delimiters. vision_tokens: flat list of codes."""
    vision_tokens = list(vision_tokens)
    n_frames = int(len(vision_tokens) / n_tokens_per_frame)
    if max_n_frames > 0 and n_frames > max_n_frames:
        idxs = np.linspace(0, n_frames - 1, max_n_frames).astype(int)
        new = []
        for idx in idxs:
            new.extend(vision_tokens[idx * n_tokens_per_frame:(idx + 1) * n_tokens_per_frame])
        vision_tokens, n_frames = new, max_n_frames
    assert n_frames > 0
    tokens = []
    for j in range(n_frames):
        tokens.extend(vision_tokens[j * n_tokens_per_frame:(j + 1) * n_tokens_per_frame])
        tokens.append(eov_token if j == n_frames - 1 else eof_token)
    return tokens


def vision_field(vision_tokens, vision_start, vision_end, **kw):
    """tokens and vision_mask of one vision field (data.py:206-219)."""
    body = frame_tokens(vision_tokens, **kw)
    tokens = list(vision_start) + body + list(vision_end)
    mask = [False] * len(vision_start) + [True] * len(body) + [False] * len(vision_end)
    return tokens, mask


def unframe_tokens(tokens, n_tokens_per_frame=256):
    """vision_generation.py:160 / :221 — drop the last token of every (P+1)-token frame."""
    t = np.asarray(tokens).reshape(-1, n_tokens_per_frame + 1)
    return t[:, :-1]
