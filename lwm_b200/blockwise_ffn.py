"""`blockwise_feedforward(cell, inputs, chunk_size, pre_remat)` — the sequence-chunked FFN wrapper the reference imports
from the ringattention package (lwm/llama.py:30, call site llama.py:728-734).

Host-side sequencing only (no kernels of its own): the FFN `cell` (any callable [B,n,D] -> [B,n,D'], e.g. the caller's
GEMM + SiLU-gate block) is applied to `chunk_size`-token slices of the sequence, so that the 4x-hidden intermediate
only ever exists for one chunk; with pre_remat=False every chunk is additionally re-materialised in the backward
(torch.utils.checkpoint), which is what the reference's `nn.remat` wrapping does when the caller has not already
done it (pre_remat=True: the cell is already re-materialising, llama.py:700-706)."""
import torch
from torch.utils.checkpoint import checkpoint


def blockwise_feedforward(cell, inputs, chunk_size, pre_remat=True):
    if inputs.dim() != 3:
        raise ValueError("blockwise_feedforward: inputs must be [batch, seq, dim]")
    S = inputs.shape[1]
    chunk_size = int(chunk_size)
    if chunk_size <= 0 or S % chunk_size:
        raise ValueError("blockwise_feedforward: chunk_size %d must divide the sequence length %d" % (chunk_size, S))
    outs = []
    for chunk in inputs.split(chunk_size, dim=1):
        if not pre_remat and torch.is_grad_enabled() and chunk.requires_grad:
            outs.append(checkpoint(cell, chunk, use_reentrant=False))
        else:
            outs.append(cell(chunk))
    return torch.cat(outs, dim=1)
