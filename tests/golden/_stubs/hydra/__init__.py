"""Stand-in for hydra (absent here); see ../README.md."""
from . import utils  # noqa: F401
