import numpy as np


def uniform(rng, shape, dtype=np.float32, minval=0.0, maxval=1.0):
    return np.random.default_rng(0).uniform(minval, maxval, size=shape).astype(dtype)
