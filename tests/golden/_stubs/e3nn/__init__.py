"""Stand-in for e3nn (absent here); see ../README.md."""
from . import o3, util  # noqa: F401
