import importlib


def _locate(path):
    parts = path.split(".")
    for i in range(len(parts), 0, -1):
        try:
            mod = importlib.import_module(".".join(parts[:i]))
        except ImportError:
            continue
        obj = mod
        for p in parts[i:]:
            obj = getattr(obj, p)
        return obj
    raise ImportError(path)


get_method = _locate
get_class = _locate

_NO_RECURSE = ("optimizer_cfg", "member_cfg", "activation_fn_cfg")


def instantiate(cfg, *args, **kwargs):
    cfg = dict(cfg)
    cfg.update(kwargs)
    target = _locate(cfg.pop("_target_"))
    cfg.pop("_recursive_", None)
    out = {}
    for k, v in cfg.items():
        if isinstance(v, dict) and "_target_" in v and k not in _NO_RECURSE:
            v = instantiate(v)
        out[k] = v
    return target(*args, **out)
