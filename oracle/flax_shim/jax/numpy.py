"""jax.numpy -> numpy (float32 arrays; every function lwm/vqgan.py uses has the same name and meaning)."""
from numpy import *  # noqa: F401,F403
from numpy import float32, int32  # noqa: F401
