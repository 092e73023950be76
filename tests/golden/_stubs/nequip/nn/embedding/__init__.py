import torch

from e3nn.o3 import Irreps
from nequip.data import AtomicDataDict
from nequip.nn import GraphModuleMixin
from oracle import nn_ref as _R


class PolynomialCutoff(torch.nn.Module):
    def __init__(self, p=6):
        super().__init__()
        self.p = float(p)

    def forward(self, x):
        return _R.polynomial_cutoff(x, self.p)


class EdgeLengthNormalizer(GraphModuleMixin, _R.EdgeLengthNormalizer):
    def __init__(self, r_max, type_names, per_edge_type_cutoff=None, irreps_in=None):
        _R.EdgeLengthNormalizer.__init__(self, r_max, type_names, per_edge_type_cutoff)
        self._init_irreps(irreps_in=irreps_in, irreps_out={AtomicDataDict.NORM_LENGTH_KEY: Irreps("1x0e")})


class BesselEdgeLengthEncoding(GraphModuleMixin, _R.BesselEdgeLengthEncoding):
    def __init__(self, cutoff, num_bessels=8, trainable=False, edge_invariant_field=AtomicDataDict.EDGE_EMBEDDING_KEY, irreps_in=None):
        _R.BesselEdgeLengthEncoding.__init__(self, num_bessels, cutoff.p, trainable)
        assert edge_invariant_field == AtomicDataDict.EDGE_EMBEDDING_KEY
        self._output_dtype = torch.get_default_dtype()
        self._init_irreps(irreps_in=irreps_in, irreps_out={edge_invariant_field: Irreps([(num_bessels, (0, 1))])})

    def forward(self, data):
        return _R.BesselEdgeLengthEncoding.forward(self, data, self._output_dtype)


class AddRadialCutoffToData(GraphModuleMixin, torch.nn.Module):
    def __init__(self, cutoff, irreps_in=None):
        super().__init__()
        self.cutoff = cutoff
        self._init_irreps(irreps_in=irreps_in, irreps_out={AtomicDataDict.EDGE_CUTOFF_KEY: Irreps("1x0e")})

    def forward(self, data):
        data[AtomicDataDict.EDGE_CUTOFF_KEY] = self.cutoff(data[AtomicDataDict.NORM_LENGTH_KEY])
        return data
