"""Config plumbing of the reference (sgm/util.py:168-185 instantiate_from_config / get_obj_from_str,
:163-166 append_dims, default/exists), restated; `sgm.` targets that this package mirrors resolve to the mirror."""
from __future__ import annotations

import importlib


def exists(x):
    return x is not None


def default(val, d):
    if exists(val):
        return val
    return d() if callable(d) and not isinstance(d, (dict, list, str)) else d


def append_dims(x, target_dims):
    """Appends trailing singleton dims until x.ndim == target_dims."""
    extra = target_dims - x.ndim
    if extra < 0:
        raise ValueError(f"input has {x.ndim} dims but target_dims is {target_dims}")
    return x[(...,) + (None,) * extra]


_MIRROR_PREFIX = "panacea_b200."


def get_obj_from_str(string: str, reload: bool = False):
    module, cls = string.rsplit(".", 1)
    if module.startswith("sgm."):
        try:
            mod = importlib.import_module(_MIRROR_PREFIX + module)
            if hasattr(mod, cls):
                return getattr(mod, cls)
        except ModuleNotFoundError:
            pass
    mod = importlib.import_module(module)
    if reload:
        importlib.reload(mod)
    return getattr(mod, cls)


def instantiate_from_config(config):
    if "target" not in config:
        if config in ("__is_first_stage__", "__is_unconditional__"):
            return None
        raise KeyError("Expected key `target` to instantiate.")
    return get_obj_from_str(config["target"])(**dict(config.get("params", dict()) or {}))
