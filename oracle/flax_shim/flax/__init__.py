from . import jax_utils, linen  # noqa: F401
