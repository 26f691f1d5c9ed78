"""Import shim: `from ringattention import ringattention` (lwm/llama.py:30) resolves to the B200 op."""
from lwm_b200.ringattention import ringattention, ringattention_inference, set_axis_group  # noqa: F401


def _not_built(name):
    def f(*a, **k):
        raise NotImplementedError("%s is outside the hot-path scope of lwm_b200 (SURVEY.md §8f next-rows)" % name)
    f.__name__ = name
    return f


blockwise_feedforward = _not_built("blockwise_feedforward")
ringattention_jax = ringattention
