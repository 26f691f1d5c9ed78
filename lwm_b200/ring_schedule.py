"""Ring schedules (placeholder — filled in with the multi-GPU path)."""


def make_plan(world, rank, Sq, Sk, causal, layout):
    raise NotImplementedError


def run_forward(*a, **k):
    raise NotImplementedError


def run_backward(*a, **k):
    raise NotImplementedError
