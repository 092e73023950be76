import functools

import torch

from nequip.utils.dtype import torch_default_dtype

_DT = {"float32": torch.float32, "float64": torch.float64}


def model_builder(fn):
    """Outermost call consumes seed / model_dtype / compile_mode (sets the RNG seed and the default
    dtype while the modules are constructed); nested builder calls pass straight through."""

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        if "model_dtype" not in kwargs and "seed" not in kwargs:
            return fn(*args, **kwargs)
        kwargs = dict(kwargs)
        seed = kwargs.pop("seed", None)
        dt = _DT[kwargs.pop("model_dtype", "float32")]
        kwargs.pop("compile_mode", None)
        if seed is not None:
            torch.manual_seed(seed)
        with torch_default_dtype(dt):
            return fn(*args, **kwargs)

    return wrapped
