import numpy as np


def resize(x, shape, method="nearest"):
    """jax.image.resize(..., 'nearest') for exact integer up-scaling: out[i, j] = in[i // s, j // s]."""
    assert method == "nearest"
    x = np.asarray(x)
    B, H, W, C = x.shape
    assert shape[0] == B and shape[3] == C and shape[1] % H == 0 and shape[2] % W == 0
    ii = np.arange(shape[1]) // (shape[1] // H)
    jj = np.arange(shape[2]) // (shape[2] // W)
    return x[:, ii][:, :, jj]
