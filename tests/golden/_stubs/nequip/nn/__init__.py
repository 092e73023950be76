"""nequip.nn look-alikes.  Graph bookkeeping (GraphModuleMixin / SequentialGraphNetwork) is
re-implemented minimally; arithmetic modules forward to the oracle's restatements."""
import torch

from e3nn.o3 import Irreps
from nequip.data import AtomicDataDict
from oracle import nn_ref as _R


# ---- graph bookkeeping -------------------------------------------------------------------
class GraphModuleMixin:
    def _init_irreps(self, irreps_in=None, my_irreps_in=None, required_irreps_in=(), irreps_out=None):
        irreps_in = {} if irreps_in is None else dict(irreps_in)
        irreps_in = {k: (None if v is None else Irreps(v)) for k, v in irreps_in.items()}
        for k in required_irreps_in:
            assert k in irreps_in, f"missing required input field {k}"
        for k, v in (my_irreps_in or {}).items():
            assert k in irreps_in and irreps_in[k] == Irreps(v), f"irreps mismatch for {k}"
        self.irreps_in = irreps_in
        out = dict(irreps_in)
        out.update({k: (None if v is None else Irreps(v)) for k, v in (irreps_out or {}).items()})
        self.irreps_out = out


class SequentialGraphNetwork(GraphModuleMixin, torch.nn.Sequential):
    def __init__(self, modules):
        super().__init__()
        mods = list(modules.items())
        for name, m in mods:
            self.add_module(name, m)
        self._init_irreps(irreps_in=mods[0][1].irreps_in, irreps_out=mods[-1][1].irreps_out)

    def forward(self, data):
        for m in self:
            data = m(data)
        return data


def replace_submodules(model, target_cls, factory):
    for name, child in list(model.named_children()):
        if type(child) is target_cls:
            setattr(model, name, factory(child))
        else:
            replace_submodules(child, target_cls, factory)
    return model


def model_modifier(persistent=False):
    def deco(fn):
        return fn

    return deco


# ---- arithmetic (oracle restatements of the published nequip modules) ---------------------
ScalarMLPFunction = _R.ScalarMLPFunction
tp_path_exists = _R.tp_path_exists


def scatter(src, index, dim=0, dim_size=None, reduce="sum"):
    assert dim == 0 and reduce == "sum"
    return _R.scatter(src, index, int(dim_size))


def with_edge_vectors_(data, with_lengths=True):
    return _R.with_edge_vectors_(data)


class ScalarMLP(GraphModuleMixin, _R.ScalarMLPFunction):
    """nequip.nn.ScalarMLP: ScalarMLPFunction applied to one field of the graph dict."""

    def __init__(self, output_dim, hidden_layers_depth=0, hidden_layers_width=None, nonlinearity="silu", bias=False,
                 forward_weight_init=True, field=AtomicDataDict.EDGE_EMBEDDING_KEY, out_field=None, irreps_in=None):
        self._init_irreps(irreps_in=irreps_in, required_irreps_in=[field])
        in_dim = self.irreps_in[field].num_irreps
        _R.ScalarMLPFunction.__init__(self, in_dim, output_dim, hidden_layers_depth, hidden_layers_width, nonlinearity, bias,
                                      forward_weight_init)
        self.field, self.out_field = field, (out_field or field)
        self.irreps_out[self.out_field] = Irreps([(output_dim, (0, 1))])

    def forward(self, data):
        data[self.out_field] = _R.ScalarMLPFunction.forward(self, data[self.field])
        return data


class AtomwiseReduce(GraphModuleMixin, torch.nn.Module):
    def __init__(self, field, out_field=None, reduce="sum", irreps_in=None):
        super().__init__()
        assert reduce == "sum"
        self.field, self.out_field = field, (out_field or f"{reduce}_{field}")
        self._init_irreps(irreps_in=irreps_in, irreps_out={self.out_field: (irreps_in or {}).get(field)})

    def forward(self, data):
        data[self.out_field] = data[self.field].sum(dim=0, keepdim=True)  # single frame
        return data


class PerTypeScaleShift(GraphModuleMixin, _R.PerTypeScaleShift):
    def __init__(self, type_names, field, out_field, scales=None, shifts=None, scales_trainable=False, shifts_trainable=False,
                 irreps_in=None):
        assert field == out_field == AtomicDataDict.PER_ATOM_ENERGY_KEY and not scales_trainable and not shifts_trainable
        _R.PerTypeScaleShift.__init__(self, len(type_names), scales, shifts)
        self._init_irreps(irreps_in=irreps_in)


class ForceStressOutput(GraphModuleMixin, torch.nn.Module):
    """forces = -dE_total/dpos (nequip.nn.ForceStressOutput without the stress branch)."""

    def __init__(self, func):
        super().__init__()
        self.model = func
        self._init_irreps(irreps_in=func.irreps_in, irreps_out=func.irreps_out)

    def forward(self, data):
        data = dict(data)
        pos = data[AtomicDataDict.POSITIONS_KEY].detach().clone().requires_grad_(True)
        data[AtomicDataDict.POSITIONS_KEY] = pos
        with torch.enable_grad():
            data = self.model(data)
            (g,) = torch.autograd.grad(data[AtomicDataDict.TOTAL_ENERGY_KEY].sum(), pos)
        data[AtomicDataDict.FORCE_KEY] = -g
        data[AtomicDataDict.POSITIONS_KEY] = pos.detach()
        return {k: (v.detach() if isinstance(v, torch.Tensor) else v) for k, v in data.items()}
