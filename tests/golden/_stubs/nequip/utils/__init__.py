import logging

from . import dtype, global_dtype  # noqa: F401


class RankedLogger(logging.LoggerAdapter):
    def __init__(self, name, rank_zero_only=True):
        super().__init__(logging.getLogger(name), {})
