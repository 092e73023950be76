from ._irreps import Irrep, Irreps  # noqa: F401
from ._spherical_harmonics import SphericalHarmonics  # noqa: F401
from ._wigner import wigner_3j  # noqa: F401
