import importlib


def instantiate(cfg, **kwargs):
    """hydra.utils.instantiate for a flat {'_target_': 'pkg.mod.Name', ...} mapping."""
    cfg = dict(cfg)
    mod, name = cfg.pop("_target_").rsplit(".", 1)
    cfg.update(kwargs)
    return getattr(importlib.import_module(mod), name)(**cfg)
