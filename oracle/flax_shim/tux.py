import inspect

from ml_collections import ConfigDict


def function_args_to_config(fn, none_arg_types=None, exclude_args=None, override_args=None):
    cfg = ConfigDict()
    for name, p in inspect.signature(fn).parameters.items():
        if name == "self" or p.default is inspect.Parameter.empty:
            continue
        cfg[name] = p.default
    return cfg


def open_file(path, mode="rb"):
    return open(path, mode)
