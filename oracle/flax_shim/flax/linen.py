"""A tiny flax.linen: Module (dataclass-style fields, @compact, setup), auto-naming of sub-modules
(`<Class>_<n>` in instantiation order per class inside a compact method; attribute names for setup()-style
modules), `self.param`, `Module.apply({'params': tree}, *args, method=...)`, and the primitives lwm/vqgan.py uses:
Conv, GroupNorm, Dropout, silu, avg_pool. Parameters are READ from the supplied tree (no initialisation)."""
import numpy as np
import torch
import torch.nn.functional as F

_STACK = []      # active (module, params-subtree, per-class counters) frames


def compact(fn):
    return fn


def silu(x):
    x = np.asarray(x, dtype=np.float32)
    return (x / (1.0 + np.exp(-x))).astype(np.float32)


def avg_pool(x, window, strides):
    raise NotImplementedError("avg_pool path is unused (resample_with_conv=True)")


class Module:
    def __init__(self, *args, **kwargs):
        fields = [k for k in getattr(type(self), "__annotations__", {})]
        for k, v in zip(fields, args):
            object.__setattr__(self, k, v)
        for k in fields[len(args):]:
            if k in kwargs:
                object.__setattr__(self, k, kwargs[k])
            elif hasattr(type(self), k):
                object.__setattr__(self, k, getattr(type(self), k))
        object.__setattr__(self, "_name", None)
        # auto-name: <Class>_<n> among the children created inside the parent's compact call
        if _STACK and _STACK[-1]["in_call"]:
            frame = _STACK[-1]
            cls = type(self).__name__
            n = frame["counters"].get(cls, 0)
            frame["counters"][cls] = n + 1
            object.__setattr__(self, "_name", "%s_%d" % (cls, n))

    # setup()-style modules name their children by attribute
    def __setattr__(self, key, value):
        if isinstance(value, Module) and _STACK and _STACK[-1].get("in_setup") and _STACK[-1]["module"] is self:
            object.__setattr__(value, "_name", key)
        object.__setattr__(self, key, value)

    def _params(self):
        return _STACK[-1]["params"]

    def param(self, name, init_fn, *a):
        return np.asarray(self._params()[name], dtype=np.float32)

    def _enter(self, params):
        frame = {"module": self, "params": params, "counters": {}, "in_call": False, "in_setup": False}
        _STACK.append(frame)
        if hasattr(self, "setup") and not getattr(self, "_setup_done", False):
            frame["in_setup"] = True
            self.setup()
            frame["in_setup"] = False
            object.__setattr__(self, "_setup_done", True)
        frame["in_call"] = True
        return frame

    def __call__(self, *args, **kwargs):
        parent = _STACK[-1]
        # parameter-less modules (Dropout, Downsample wrappers without own params ...) have no entry in the tree
        sub = parent["params"].get(self._name, {}) if self._name is not None else parent["params"]
        self._enter(sub)
        try:
            return type(self).__call_impl__(self, *args, **kwargs)
        finally:
            _STACK.pop()

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        if "__call__" in cls.__dict__:
            cls.__call_impl__ = cls.__dict__["__call__"]
            del cls.__call__

    def apply(self, variables, *args, method=None, **kwargs):
        self._enter(variables["params"])
        try:
            fn = method if method is not None else (lambda *a, **k: type(self).__call_impl__(self, *a, **k))
            return fn(*args, **kwargs)
        finally:
            _STACK.pop()


class Conv(Module):
    features: int
    kernel_size: object
    strides: object = None
    padding: object = "SAME"

    def __call_impl__(self, x):
        p = self._params()
        w = torch.as_tensor(np.asarray(p["kernel"], dtype=np.float32))      # HWIO
        b = torch.as_tensor(np.asarray(p["bias"], dtype=np.float32))
        k = w.shape[0]
        s = 1 if self.strides is None else int(self.strides[0])
        pad = (k // 2) if self.padding == "SAME" else 0
        assert self.padding in ("SAME", "VALID") and w.shape[3] == self.features
        xt = torch.as_tensor(np.asarray(x, dtype=np.float32)).permute(0, 3, 1, 2)
        y = F.conv2d(xt, w.permute(3, 2, 0, 1), bias=None, stride=s, padding=pad).permute(0, 2, 3, 1) + b
        return y.numpy()


class GroupNorm(Module):
    num_groups: int = 32
    epsilon: float = 1e-6

    def __call_impl__(self, x):
        p = self._params()
        x = np.asarray(x, dtype=np.float32)
        N, H, W, C = x.shape
        g = self.num_groups
        xg = x.reshape(N, H * W, g, C // g)
        mean = xg.mean(axis=(1, 3), keepdims=True, dtype=np.float32)
        msq = (xg * xg).mean(axis=(1, 3), keepdims=True, dtype=np.float32)
        var = np.maximum(msq - mean * mean, 0.0)
        y = (xg - mean) / np.sqrt(var + np.float32(self.epsilon))
        return (y.reshape(N, H, W, C) * np.asarray(p["scale"], np.float32) + np.asarray(p["bias"], np.float32)).astype(
            np.float32)


class Dropout(Module):
    rate: float = 0.0
    deterministic: bool = True

    def __call_impl__(self, x):
        assert self.deterministic or self.rate == 0.0
        return x
