import torch

from oracle import o3_ref

from ._irreps import Irreps


class SphericalHarmonics(torch.nn.Module):
    """e3nn.o3.SphericalHarmonics(irreps_out, normalize, normalization) for irreps 0..lmax."""

    def __init__(self, irreps_out, normalize, normalization="integral", irreps_in=None):
        super().__init__()
        if isinstance(irreps_out, int):
            irreps_out = Irreps.spherical_harmonics(irreps_out)
        self.irreps_out = Irreps(irreps_out)
        ls = [mi.ir.l for mi in self.irreps_out]
        assert ls == list(range(len(ls))) and all(mi.mul == 1 and mi.ir.p == (-1) ** mi.ir.l for mi in self.irreps_out)
        self.lmax, self.normalize, self.normalization = len(ls) - 1, normalize, normalization

    def forward(self, x):
        return o3_ref.spherical_harmonics(self.lmax, x, self.normalize, self.normalization)
