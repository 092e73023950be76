"""Oracle modules (test infrastructure; see oracle/__init__.py).

Plain-PyTorch restatement, one class per reference class, following the cited
lines in behaviour.  ``data`` is a plain dict keyed by the nequip
AtomicDataDict strings.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .o3_ref import Irrep, Irreps, spherical_harmonics, wigner_3j

# AtomicDataDict keys (nequip.data.AtomicDataDict; strings recalled)
POSITIONS_KEY = "pos"
EDGE_INDEX_KEY = "edge_index"
ATOM_TYPE_KEY = "atom_types"
CELL_KEY = "cell"
EDGE_CELL_SHIFT_KEY = "edge_cell_shift"
EDGE_VECTORS_KEY = "edge_vectors"
EDGE_LENGTH_KEY = "edge_lengths"
NORM_LENGTH_KEY = "normed_edge_lengths"
EDGE_TYPE_KEY = "edge_type"
EDGE_ATTRS_KEY = "edge_attrs"
EDGE_EMBEDDING_KEY = "edge_embedding"
EDGE_FEATURES_KEY = "edge_features"
EDGE_ENERGY_KEY = "edge_energy"
PER_ATOM_ENERGY_KEY = "atomic_energy"
TOTAL_ENERGY_KEY = "total_energy"
FORCE_KEY = "forces"
STRESS_KEY = "stress"
VIRIAL_KEY = "virial"


def silu_second_moment_gain() -> float:
    """1/sqrt(E_{z~N(0,1)}[silu(z)^2]) -- e3nn ``normalize2mom`` constant (e3nn estimates it
    by sampling; we integrate).  SURVEY appendix A.3."""
    z = np.linspace(-12.0, 12.0, 240001)
    w = np.exp(-0.5 * z * z) / math.sqrt(2 * math.pi)
    s = z / (1.0 + np.exp(-z))
    return float(1.0 / math.sqrt(np.trapezoid(s * s * w, z)))


_SILU_GAIN = silu_second_moment_gain()


def scatter(src, index, dim_size: int):
    """nequip.nn.scatter(reduce='sum', dim=0): zero-initialised segment sum."""
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add_(0, index, src)


def tp_path_exists(irreps_a, irreps_b, ir_out) -> bool:
    ir_out = Irrep(ir_out)
    return any(ir_out in (a * b) for _, a in Irreps(irreps_a) for _, b in Irreps(irreps_b))


class ScalarMLPFunction(torch.nn.Module):
    """nequip.nn.ScalarMLPFunction: x @ (alpha_k W_k), SiLU between."""

    def __init__(
        self,
        input_dim: int,
        output_dim: int,
        hidden_layers_depth: int = 0,
        hidden_layers_width: Optional[int] = None,
        nonlinearity: Optional[str] = "silu",
        bias: bool = False,
        forward_weight_init: bool = True,
    ):
        super().__init__()
        assert not bias
        assert nonlinearity in ("silu", None)
        self.dims = [input_dim] + hidden_layers_depth * [hidden_layers_width] + [output_dim]
        self.nonlinearity = nonlinearity
        self.is_nonlinear = hidden_layers_depth > 0 and nonlinearity is not None
        self.weights = torch.nn.ParameterList()
        self.alphas: List[float] = []
        gain = 1.0
        for h_in, h_out in zip(self.dims, self.dims[1:]):
            w = torch.empty(h_in, h_out)
            torch.nn.init.uniform_(w, -math.sqrt(3), math.sqrt(3))
            self.weights.append(torch.nn.Parameter(w))
            norm_dim = h_in if forward_weight_init else h_out
            self.alphas.append(gain / math.sqrt(norm_dim))
            gain = _SILU_GAIN if nonlinearity == "silu" else 1.0

    def forward(self, x):
        n = len(self.weights)
        for k, (w, a) in enumerate(zip(self.weights, self.alphas)):
            x = x @ (a * w)
            if k < n - 1 and self.nonlinearity == "silu":
                x = torch.nn.functional.silu(x)
        return x


class MakeWeightedChannels(torch.nn.Module):
    """allegro/nn/_strided/_channels.py:7-63."""

    def __init__(self, irreps_in, multiplicity_out: int, alpha: float = 1.0, weight_individual_irreps: bool = True):
        super().__init__()
        irreps_in = Irreps(irreps_in)
        assert all(mul == 1 for mul, _ in irreps_in)
        self._num_irreps = len(irreps_in)
        self.multiplicity_out = multiplicity_out
        self.weight_individual_irreps = weight_individual_irreps
        self.alpha = alpha
        if not weight_individual_irreps:
            self.weight_numel = multiplicity_out
            self.register_buffer("_rtoi", torch.Tensor())  # persistent, empty (_channels.py:31)
            return
        self.weight_numel = len(irreps_in) * multiplicity_out
        rtoi = torch.zeros(self._num_irreps, irreps_in.dim)
        for i, sl in enumerate(irreps_in.slices()):
            rtoi[i, sl] = alpha
        self.register_buffer("_rtoi", rtoi, persistent=False)

    def forward(self, edge_attr, weights):
        if self.weight_individual_irreps:
            aux = torch.mm(weights.reshape(-1, self._num_irreps), self._rtoi.to(weights.dtype)).view(
                edge_attr.size(0), self.multiplicity_out, self._rtoi.shape[1]
            )
            return edge_attr.unsqueeze(1) * aux
        return weights.unsqueeze(-1) * (self.alpha * edge_attr.unsqueeze(-2))


class Contracter(torch.nn.Module):
    """allegro/nn/_strided/_contract.py:11-251 (table build :80-168, forward :185-211,
    _contract :213-251).  ``chunk`` only bounds the size of the dense intermediate; the
    arithmetic is the reference's."""

    def __init__(
        self,
        irreps_in1,
        irreps_in2,
        irreps_out,
        mul: int,
        instructions=None,
        path_channel_coupling: bool = True,
        scatter_factor: Optional[float] = None,
        irrep_normalization: Optional[str] = "component",
        chunk: int = 4096,
    ):
        super().__init__()
        self.scatter_factor = scatter_factor
        self.irreps_in1, self.irreps_in2, self.irreps_out = Irreps(irreps_in1), Irreps(irreps_in2), Irreps(irreps_out)
        for irr in (self.irreps_in1, self.irreps_in2, self.irreps_out):
            assert all(m == 1 for m, _ in irr)
        self.instructions = instructions
        if instructions is None:
            instructions = [
                (i1, i2, io)
                for io, (_, iro) in enumerate(self.irreps_out)
                for i1, (_, ir1) in enumerate(self.irreps_in1)
                for i2, (_, ir2) in enumerate(self.irreps_in2)
                if iro in ir1 * ir2
            ]
        self.mul = mul
        self.base_dim1, self.base_dim2, self.base_dim_out = self.irreps_in1.dim, self.irreps_in2.dim, self.irreps_out.dim
        self.num_paths = len(instructions)
        assert self.num_paths > 0
        self.irrep_normalization = irrep_normalization
        self.chunk = chunk
        idx_list, val_list = [], []
        s1, s2, so = self.irreps_in1.slices(), self.irreps_in2.slices(), self.irreps_out.slices()
        for i1, i2, io in instructions:
            ir1, ir2, iro = self.irreps_in1[i1][1], self.irreps_in2[i2][1], self.irreps_out[io][1]
            assert ir1.p * ir2.p == iro.p and abs(ir1.l - ir2.l) <= iro.l <= ir1.l + ir2.l
            w = torch.from_numpy(np.array(wigner_3j(ir1.l, ir2.l, iro.l))).to(torch.get_default_dtype())
            nz = w.nonzero()
            vals = w[nz[:, 0], nz[:, 1], nz[:, 2]].clone()
            if irrep_normalization == "component":
                vals *= math.sqrt(2 * iro.l + 1)
            else:
                assert irrep_normalization is None
            nz = nz + torch.tensor([s1[i1].start, s2[i2].start, so[io].start])
            idx_list.append(nz)
            val_list.append(vals)
        self.w3j_is_ij_diagonal = (self.base_dim1 == self.base_dim2) and all(
            bool(torch.all(e[:, 0] == e[:, 1])) for e in idx_list
        )
        if self.w3j_is_ij_diagonal:
            w3j = torch.zeros(self.num_paths, self.base_dim1, self.base_dim_out)
            for p, (ix, v) in enumerate(zip(idx_list, val_list)):
                w3j[p, ix[:, 0], ix[:, 2]] = v
        else:
            w3j = torch.zeros(self.num_paths, self.base_dim1, self.base_dim2, self.base_dim_out)
            for p, (ix, v) in enumerate(zip(idx_list, val_list)):
                w3j[p, ix[:, 0], ix[:, 1], ix[:, 2]] = v
        if self.num_paths == 1:
            w3j = w3j.squeeze(0)
        self.register_buffer("w3j", w3j)
        self.path_channel_coupling = path_channel_coupling
        shape = (mul,) if path_channel_coupling else tuple()
        if self.num_paths > 1:
            shape = shape + (self.num_paths,)
        self.weights = torch.nn.Parameter(torch.empty(shape).uniform_(-math.sqrt(3), math.sqrt(3)))
        ij = "i" if self.w3j_is_ij_diagonal else "ij"
        p = "p" if self.num_paths > 1 else ""
        u = "u" if path_channel_coupling else ""
        self._weight_w3j_einstr = f"{u}{p},{p}{ij}k->{u}{ij}k"

    def forward(self, x1, x2, idxs, scatter_dim_size):
        if self.scatter_factor is not None:
            x2 = self.scatter_factor * x2
        n = int(scatter_dim_size)
        x2 = torch.index_select(scatter(x2, idxs, n), 0, idxs)
        x1 = x1.reshape(-1, self.mul, self.base_dim1)
        x2 = x2.reshape(-1, self.mul, self.base_dim2)
        return self._contract(x1, x2)

    def _contract(self, x1, x2):
        ww3j = torch.einsum(self._weight_w3j_einstr, self.weights, self.w3j)
        outs = []
        for a in range(0, x1.shape[0], self.chunk) if x1.shape[0] else [0]:
            c1, c2 = x1[a : a + self.chunk], x2[a : a + self.chunk]
            if self.w3j_is_ij_diagonal:
                outer = c1 * c2
                if self.path_channel_coupling:
                    out = torch.sum(outer.unsqueeze(-1) * ww3j, 2)
                else:
                    out = torch.mm(outer.reshape(outer.size(0) * outer.size(1), outer.size(2)), ww3j).view(outer.size(0), outer.size(1), ww3j.size(1))
            else:
                outer = c1.unsqueeze(-1) * c2.unsqueeze(-2)
                if self.path_channel_coupling:
                    out = torch.sum(outer.unsqueeze(-1) * ww3j, (2, 3))
                else:
                    out = torch.mm(
                        outer.reshape(outer.size(0) * outer.size(1), outer.size(2) * outer.size(3)), ww3j.reshape(-1, ww3j.size(2))
                    ).view(-1, self.mul, ww3j.size(2))
            outs.append(out)
        return torch.cat(outs, 0)


def with_edge_vectors_(data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """nequip with_edge_vectors_: r_ij = pos[j] - pos[i] (+ shift @ cell)."""
    if EDGE_VECTORS_KEY in data:
        return data
    pos, ei = data[POSITIONS_KEY], data[EDGE_INDEX_KEY]
    vec = pos[ei[1]] - pos[ei[0]]
    if EDGE_CELL_SHIFT_KEY in data and CELL_KEY in data:
        vec = vec + data[EDGE_CELL_SHIFT_KEY].to(pos.dtype) @ data[CELL_KEY].view(3, 3).to(pos.dtype)
    data[EDGE_VECTORS_KEY] = vec
    data[EDGE_LENGTH_KEY] = vec.norm(dim=-1)
    return data


class EdgeLengthNormalizer(torch.nn.Module):
    """nequip EdgeLengthNormalizer: x = r / r_max (per-edge-type table if given)."""

    def __init__(self, r_max: float, type_names: Sequence[str], per_edge_type_cutoff=None):
        super().__init__()
        self.r_max = float(r_max)
        self.num_types = len(type_names)
        self._per_type = per_edge_type_cutoff is not None
        if self._per_type:
            tab = torch.full((self.num_types, self.num_types), float(r_max), dtype=torch.float64)
            for a, va in per_edge_type_cutoff.items():
                ia = list(type_names).index(a)
                if isinstance(va, dict):
                    for b, vb in va.items():
                        tab[ia, list(type_names).index(b)] = float(vb)
                else:
                    tab[ia, :] = float(va)
            self.register_buffer("rmax_table", tab)

    def forward(self, data):
        data = with_edge_vectors_(data)
        r = data[EDGE_LENGTH_KEY]
        ei = data[EDGE_INDEX_KEY]
        et = data[ATOM_TYPE_KEY].reshape(-1)[ei]  # [2,E]
        data[EDGE_TYPE_KEY] = et
        if self._per_type:
            rmax = self.rmax_table[et[0], et[1]].to(r.dtype)
            data[NORM_LENGTH_KEY] = (r / rmax).unsqueeze(-1)
        else:
            data[NORM_LENGTH_KEY] = (r / self.r_max).unsqueeze(-1)
        return data


def polynomial_cutoff(x, p: float = 6.0):
    """nequip PolynomialCutoff."""
    out = 1.0 - ((p + 1.0) * (p + 2.0) / 2.0) * x**p + p * (p + 2.0) * x ** (p + 1.0) - (p * (p + 1.0) / 2.0) * x ** (p + 2.0)
    return out * (x < 1.0)


class BesselEdgeLengthEncoding(torch.nn.Module):
    """nequip BesselEdgeLengthEncoding: b_n(x) = sin(n pi x)/(pi x) * cutoff(x), n=1..num_bessels."""

    def __init__(self, num_bessels: int = 8, polynomial_cutoff_p: float = 6.0, trainable: bool = False):
        super().__init__()
        self.p = float(polynomial_cutoff_p)
        w = torch.linspace(1.0, num_bessels, num_bessels, dtype=torch.float64).unsqueeze(0)
        if trainable:
            self.bessel_weights = torch.nn.Parameter(w)
        else:
            self.register_buffer("bessel_weights", w)

    def forward(self, data, model_dtype):
        x = data[NORM_LENGTH_KEY]  # [E,1]
        bw = self.bessel_weights.to(x.dtype)
        bessel = torch.sinc(x * bw) * bw
        data[EDGE_EMBEDDING_KEY] = (bessel * polynomial_cutoff(x, self.p)).to(model_dtype)
        return data


class ProductTypeEmbedding(torch.nn.Module):
    """allegro/nn/_edgeembed.py:13-85."""

    def __init__(self, num_types: int, in_dim: int, initial_embedding_dim: int, forward_weight_init: bool = True):
        super().__init__()
        assert initial_embedding_dim % 2 == 0
        self.center_embed = torch.nn.Embedding(num_types, initial_embedding_dim // 2)
        self.neighbor_embed = torch.nn.Embedding(num_types, initial_embedding_dim // 2)
        self.basis_linear = ScalarMLPFunction(in_dim, initial_embedding_dim, forward_weight_init=forward_weight_init)

    def forward(self, data):
        et = data[EDGE_TYPE_KEY]
        type_embed = torch.cat((self.center_embed(et[0]), self.neighbor_embed(et[1])), dim=-1)
        data[EDGE_EMBEDDING_KEY] = type_embed * self.basis_linear(data[EDGE_EMBEDDING_KEY])
        return data


class TwoBodyBesselScalarEmbed(torch.nn.Module):
    """allegro/nn/scalarembed.py:19-81 (bessel_encode -> type_embed)."""

    def __init__(self, type_names, num_bessels=8, bessel_trainable=False, polynomial_cutoff_p=6, module_output_dim=64, forward_weight_init=True):
        super().__init__()
        self.bessel_encode = BesselEdgeLengthEncoding(num_bessels, polynomial_cutoff_p, bessel_trainable)
        self.type_embed = ProductTypeEmbedding(len(type_names), num_bessels, module_output_dim, forward_weight_init)
        self.out_dim = module_output_dim

    def forward(self, data, model_dtype):
        return self.type_embed(self.bessel_encode(data, model_dtype))


class PerClassSpline(torch.nn.Module):
    """allegro/nn/spline.py:8-89: per-class weighted sum of finite-support cos^2-type bumps on [0, 1];
    every bump (hence the embedding and its derivative) vanishes at x = 1."""

    def __init__(self, num_classes: int, num_channels: int, num_splines: int, spline_span: int, dtype=torch.float64):
        super().__init__()
        assert 0 <= spline_span <= num_splines and num_splines > 0
        self.num_classes, self.num_channels, self.num_splines, self.spline_span = num_classes, num_channels, num_splines, spline_span
        lower = torch.arange(-spline_span, num_splines - spline_span, dtype=dtype) / num_splines
        diff = (spline_span + 1) / num_splines
        self.register_buffer("lower", lower)
        self.register_buffer("upper", lower + diff)
        self._const = 2 * math.pi / diff
        self.class_embed = torch.nn.Embedding(num_classes, num_channels * num_splines, dtype=dtype)

    def _get_basis(self, x):
        t = self._const * (torch.clamp(x, min=self.lower, max=self.upper) - self.lower)
        return 0.25 * (1 - torch.cos(t)).square()

    def forward(self, x, classes):
        w = self.class_embed(classes).view(classes.size(0), self.num_channels, self.num_splines)
        return torch.bmm(w, self._get_basis(x).unsqueeze(-1)).squeeze(-1)


class TwoBodySplineScalarEmbed(torch.nn.Module):
    """allegro/nn/scalarembed.py:84-175 (weights and evaluation in the global dtype fp64, output cast to the model dtype)."""

    def __init__(self, type_names, num_splines=16, spline_span=12, module_output_dim=64, forward_weight_init=True):
        super().__init__()
        self.num_types = len(type_names)
        self.spline = PerClassSpline(self.num_types * self.num_types, module_output_dim, num_splines, spline_span, dtype=torch.float64)
        bound = math.sqrt(3 / spline_span) if forward_weight_init else math.sqrt(3 / module_output_dim)
        torch.nn.init.uniform_(self.spline.class_embed.weight, a=-bound, b=bound)
        self.out_dim = module_output_dim

    def forward(self, data, model_dtype):
        et = data[EDGE_TYPE_KEY]
        data[EDGE_EMBEDDING_KEY] = self.spline(data[NORM_LENGTH_KEY], et[0] * self.num_types + et[1]).to(model_dtype)
        return data


class TwoBodySphericalHarmonicTensorEmbed(torch.nn.Module):
    """allegro/nn/tensorembed.py:16-96."""

    def __init__(self, lmax: int, num_tensor_features: int, scalar_dim: int, forward_weight_init=True, weight_individual_irreps=True):
        super().__init__()
        self.lmax = lmax
        irreps = Irreps.spherical_harmonics(lmax)
        self._edge_weighter = MakeWeightedChannels(irreps, num_tensor_features, weight_individual_irreps=weight_individual_irreps)
        self.env_embed_linear = ScalarMLPFunction(scalar_dim, self._edge_weighter.weight_numel, forward_weight_init=forward_weight_init)
        self._output_dtype = torch.get_default_dtype()

    def forward(self, data):
        data = with_edge_vectors_(data)
        weights = self.env_embed_linear(data[EDGE_EMBEDDING_KEY])
        edge_sh = spherical_harmonics(self.lmax, data[EDGE_VECTORS_KEY]).to(self._output_dtype)
        data[EDGE_ATTRS_KEY] = edge_sh
        data[EDGE_FEATURES_KEY] = self._edge_weighter(edge_sh, weights)
        return data


def allegro_layer_irreps(input_irreps: Irreps, allowed: Irreps, num_layers: int):
    """The forward build + backward pruning of allegro/nn/_allegro.py:101-160.
    Returns (tps_irreps_in, tps_irreps_out), env irreps = input_irreps."""
    env = Irreps([(1, ir) for _, ir in input_irreps])
    arg = env
    tps = [arg]
    for layer in range(num_layers):
        ir_out = Irreps([(1, (0, 1))]) if layer == num_layers - 1 else allowed
        ir_out = Irreps([(mul, ir) for mul, ir in ir_out if tp_path_exists(arg, env, ir)])
        arg = ir_out
        tps.append(ir_out)
    out = tps[-1]
    new = [out]
    for arg in reversed(tps[:-1]):
        keep = []
        for mul, arg_ir in arg:
            for _, env_ir in env:
                if any(i in out for i in arg_ir * env_ir):
                    keep.append((mul, arg_ir))
                    break
        keep = Irreps(keep)
        new.append(keep)
        out = keep
    tps = list(reversed(new))
    assert tps[-1].lmax == 0
    return tps[:-1], tps[1:]


class Allegro_Module(torch.nn.Module):
    """allegro/nn/_allegro.py:17-301."""

    def __init__(
        self,
        num_layers: int,
        num_scalar_features: int,
        num_tensor_features: int,
        tensor_track_allowed_irreps,
        input_irreps,
        scalar_input_dim: int,
        avg_num_neighbors: float,
        tp_path_channel_coupling: bool = True,
        weight_individual_irreps: bool = True,
        latent_kwargs: Optional[dict] = None,
    ):
        super().__init__()
        latent_kwargs = dict(latent_kwargs or {})
        assert num_layers >= 1 and avg_num_neighbors is not None
        self.num_layers, self.num_scalar_features, self.num_tensor_features = num_layers, num_scalar_features, num_tensor_features
        input_irreps = Irreps(input_irreps)
        allowed = Irreps(tensor_track_allowed_irreps)
        self._env_weighter = MakeWeightedChannels(input_irreps, num_tensor_features, weight_individual_irreps=weight_individual_irreps)
        self.first_layer_env_embed_projection = ScalarMLPFunction(
            scalar_input_dim, num_scalar_features + self._env_weighter.weight_numel
        )
        env = Irreps([(1, ir) for _, ir in input_irreps])
        assert env[0][1] == Irrep("0e")
        ins, outs = allegro_layer_irreps(input_irreps, allowed, num_layers)
        self.tps_irreps_in, self.tps_irreps_out = ins, outs
        self.latents = torch.nn.ModuleList()
        self.tps = torch.nn.ModuleList()
        self._n_scalar_outs = []
        for layer, (arg, out) in enumerate(zip(ins, outs)):
            tp = Contracter(
                Irreps([(1, ir) for _, ir in arg]),
                env,
                Irreps([(1, ir) for _, ir in out]),
                mul=num_tensor_features,
                path_channel_coupling=tp_path_channel_coupling,
                scatter_factor=1.0 / math.sqrt(avg_num_neighbors),
            )
            self.tps.append(tp)
            self._n_scalar_outs.append(1)
            assert tp.irreps_out[0][1] == Irrep("0e")
            self.latents.append(
                ScalarMLPFunction(
                    input_dim=num_scalar_features * (layer + 1) + num_tensor_features,
                    output_dim=num_scalar_features + (self._env_weighter.weight_numel if layer < num_layers - 1 else 0),
                    **latent_kwargs,
                )
            )

    def forward(self, data):
        edge_center = data[EDGE_INDEX_KEY][0]
        num_atoms = data[POSITIONS_KEY].shape[0]
        tensor_basis, tensor_features = data[EDGE_ATTRS_KEY], data[EDGE_FEATURES_KEY]
        S, wn = self.num_scalar_features, self._env_weighter.weight_numel
        projection = self.first_layer_env_embed_projection(data[EDGE_EMBEDDING_KEY])
        acc = [projection.narrow(-1, 0, S)]
        env_w = projection.narrow(-1, S, wn)
        for layer, (latent, tp) in enumerate(zip(self.latents, self.tps)):
            env_w_edges = self._env_weighter(tensor_basis, env_w)
            tensor_features = tp(tensor_features, env_w_edges, edge_center, num_atoms)
            scalars = tensor_features[:, :, :1].reshape(tensor_features.shape[0], tensor_features.shape[1])  # explicit sizes: E may be 0
            latents = latent(torch.cat(acc + [scalars], dim=-1))
            acc.append(latents.narrow(-1, 0, S))
            if layer < self.num_layers - 1:
                env_w = latents.narrow(-1, S, wn)
        data[EDGE_FEATURES_KEY] = torch.cat(acc, dim=-1)
        return data


class EdgewiseReduce(torch.nn.Module):
    """allegro/nn/edgewise.py:10-60 (reduce='sum')."""

    def __init__(self, field: str, out_field: str, factor: Optional[float] = None):
        super().__init__()
        self.field, self.out_field, self._factor = field, out_field, factor

    def forward(self, data):
        edge_data = data[self.field]
        if self._factor is not None:
            edge_data = edge_data * self._factor
        data[self.out_field] = scatter(edge_data, data[EDGE_INDEX_KEY][0], data[POSITIONS_KEY].shape[0])
        return data


class PerTypeScaleShift(torch.nn.Module):
    """nequip PerTypeScaleShift: E_i * scale[t_i] + shift[t_i]."""

    def __init__(self, num_types: int, scales=None, shifts=None):
        super().__init__()

        def _tab(v, default):
            if v is None:
                return torch.full((num_types,), default, dtype=torch.float64)
            t = torch.as_tensor(v, dtype=torch.float64).reshape(-1)
            return t.expand(num_types).clone() if t.numel() == 1 else t

        self.register_buffer("scales", _tab(scales, 1.0))
        self.register_buffer("shifts", _tab(shifts, 0.0))

    def forward(self, data):
        t = data[ATOM_TYPE_KEY].reshape(-1)
        e = data[PER_ATOM_ENERGY_KEY]
        data[PER_ATOM_ENERGY_KEY] = e * self.scales[t].to(e.dtype).unsqueeze(-1) + self.shifts[t].to(e.dtype).unsqueeze(-1)
        return data


# --------------------------------------------------------------------------------------
# ZBL pair potential (nequip.nn.pair_potential.ZBL; call site allegro/model/allegro_models.py:270-288)
# --------------------------------------------------------------------------------------
# PARITY UNPINNED: nequip is not vendored under /root/reference and not installable here, so there is no reference code or
# golden vector for this module.  Restated from its published algorithm: LAMMPS pair_style zbl with the constants of
# pair_zbl_const.h, multiplied by the polynomial cutoff the Allegro builder puts in front (PolynomialCutoff(6) on the
# normalised edge length) and by 1/2 (every pair appears as two directed edges).  Checked in tests against an independent
# closed-form evaluation on dimers and by finite differences.
_CHEMICAL_SYMBOLS = (
    "X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd "
    "In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn Fr Ra Ac Th Pa U Np Pu "
    "Am Cm Bk Cf Es Fm Md No Lr"
).split()
ATOMIC_NUMBERS = {sym: z for z, sym in enumerate(_CHEMICAL_SYMBOLS)}
_QQR2E = {"metal": 14.399645, "real": 332.06371}  # LAMMPS force->qqr2e


class ZBL(torch.nn.Module):
    def __init__(self, type_names: Sequence[str], chemical_species: Optional[Sequence[str]] = None, units: str = "metal", cutoff_p: float = 6.0):
        super().__init__()
        species = list(chemical_species) if chemical_species is not None else list(type_names)
        assert len(species) == len(type_names)
        self.register_buffer("atomic_numbers", torch.tensor([float(ATOMIC_NUMBERS[s]) for s in species], dtype=torch.float64))
        self.qqr2exesquare = _QQR2E[units] * 0.5
        self.cutoff_p = float(cutoff_p)

    def forward(self, data):
        r = data[EDGE_LENGTH_KEY]
        ei = data[EDGE_INDEX_KEY]
        et = data[EDGE_TYPE_KEY]
        Z = self.atomic_numbers.to(r.dtype)
        zi, zj = Z[et[0]], Z[et[1]]
        x = (zi.pow(0.23) + zj.pow(0.23)) * r / 0.46850
        psi = 0.02817 * torch.exp(-0.20162 * x) + 0.28022 * torch.exp(-0.40290 * x) + 0.50986 * torch.exp(-0.94229 * x) + 0.18175 * torch.exp(-3.19980 * x)
        eng = self.qqr2exesquare * (zi * zj / r) * psi * polynomial_cutoff(data[NORM_LENGTH_KEY].squeeze(-1).to(r.dtype), self.cutoff_p)
        n = data[PER_ATOM_ENERGY_KEY].shape[0]
        atomic = torch.zeros(n, dtype=eng.dtype, device=eng.device).index_add_(0, ei[0], eng)
        data[PER_ATOM_ENERGY_KEY] = data[PER_ATOM_ENERGY_KEY] + atomic.unsqueeze(-1).to(data[PER_ATOM_ENERGY_KEY].dtype)
        return data
