from . import jit  # noqa: F401
