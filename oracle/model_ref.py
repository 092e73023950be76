"""Oracle model builders (test infrastructure; see oracle/__init__.py).

Restates allegro/model/allegro_models.py:70-300 (AllegroEnergyModel / AllegroModel /
FullAllegroEnergyModel) and nequip's ForceStressOutput (forces = -dE/dpos by autograd).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import torch

from . import nn_ref as R
from .o3_ref import Irreps


class _default_dtype:
    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        self.prev = torch.get_default_dtype()
        torch.set_default_dtype(self.dtype)

    def __exit__(self, *a):
        torch.set_default_dtype(self.prev)


_DTYPES = {"float32": torch.float32, "float64": torch.float64}


class AllegroEnergyOracle(torch.nn.Module):
    """FullAllegroEnergyModel (allegro_models.py:112-300) as one module; sub-module names are
    the SequentialGraphNetwork keys (:222-228,262-268,297) so state_dict prefixes line up."""

    def __init__(
        self,
        r_max: float,
        type_names: Sequence[str],
        l_max: int,
        parity: bool = True,
        radial_chemical_embed: Optional[Dict] = None,
        radial_chemical_embed_dim: Optional[int] = None,
        per_edge_type_cutoff=None,
        scalar_embed_mlp_hidden_layers_depth: int = 1,
        scalar_embed_mlp_hidden_layers_width: int = 64,
        scalar_embed_mlp_nonlinearity: Optional[str] = "silu",
        num_layers: int = 2,
        num_scalar_features: int = 64,
        num_tensor_features: int = 16,
        allegro_mlp_hidden_layers_depth: int = 1,
        allegro_mlp_hidden_layers_width: int = 64,
        allegro_mlp_nonlinearity: Optional[str] = "silu",
        tp_path_channel_coupling: bool = True,
        readout_mlp_hidden_layers_depth: int = 1,
        readout_mlp_hidden_layers_width: int = 32,
        readout_mlp_nonlinearity: Optional[str] = "silu",
        avg_num_neighbors: Optional[float] = None,
        weight_individual_irreps: bool = True,
        per_type_energy_scales=None,
        per_type_energy_shifts=None,
        pair_potential: Optional[Dict] = None,
        forward_normalize: bool = True,
        seed: int = 0,
        model_dtype: str = "float32",
    ):
        super().__init__()
        self.model_dtype = _DTYPES[model_dtype]
        torch.manual_seed(seed)
        with _default_dtype(self.model_dtype):
            # irreps (allegro_models.py:76-86)
            irreps_edge_sh = Irreps.spherical_harmonics(l_max, p=-1)
            if parity:
                allowed = Irreps([(1, (l, p)) for l in range(l_max + 1) for p in (1, -1)])
            else:
                allowed = irreps_edge_sh
            rc = dict(radial_chemical_embed or {})
            target = rc.pop("_target_", "allegro.nn.TwoBodyBesselScalarEmbed").rsplit(".", 1)[-1]
            embed_cls = {"TwoBodyBesselScalarEmbed": R.TwoBodyBesselScalarEmbed, "TwoBodySplineScalarEmbed": R.TwoBodySplineScalarEmbed}[target]
            S = num_scalar_features
            self.edge_norm = R.EdgeLengthNormalizer(r_max, type_names, per_edge_type_cutoff)
            self.radial_chemical_embed = embed_cls(
                type_names=type_names,
                module_output_dim=S if radial_chemical_embed_dim is None else radial_chemical_embed_dim,
                forward_weight_init=forward_normalize,
                **rc,
            )
            self.scalar_embed_mlp = R.ScalarMLPFunction(
                self.radial_chemical_embed.out_dim,
                S,
                scalar_embed_mlp_hidden_layers_depth,
                scalar_embed_mlp_hidden_layers_width,
                scalar_embed_mlp_nonlinearity,
                forward_weight_init=forward_normalize,
            )
            # NOTE the reference builder does not forward `weight_individual_irreps` to the tensor embedding
            # (allegro_models.py:185-193): the initial features always use per-irrep weights.
            self.tensor_embed = R.TwoBodySphericalHarmonicTensorEmbed(
                l_max, num_tensor_features, S, forward_weight_init=forward_normalize,
            )
            self.allegro = R.Allegro_Module(
                num_layers=num_layers,
                num_scalar_features=S,
                num_tensor_features=num_tensor_features,
                tensor_track_allowed_irreps=allowed,
                input_irreps=irreps_edge_sh,
                scalar_input_dim=S,
                avg_num_neighbors=avg_num_neighbors,
                tp_path_channel_coupling=tp_path_channel_coupling,
                weight_individual_irreps=weight_individual_irreps,
                latent_kwargs=dict(
                    hidden_layers_depth=allegro_mlp_hidden_layers_depth,
                    hidden_layers_width=allegro_mlp_hidden_layers_width,
                    nonlinearity=allegro_mlp_nonlinearity,
                    bias=False,
                    forward_weight_init=forward_normalize,
                ),
            )
            self.edge_readout = R.ScalarMLPFunction(
                S * (num_layers + 1),
                1,
                readout_mlp_hidden_layers_depth,
                readout_mlp_hidden_layers_width,
                readout_mlp_nonlinearity,
                forward_weight_init=forward_normalize,
            )
            self.edge_eng_sum = R.EdgewiseReduce(
                R.EDGE_ENERGY_KEY, R.PER_ATOM_ENERGY_KEY, factor=1.0 / math.sqrt(2 * avg_num_neighbors)
            )
            self.per_type_energy_scale_shift = R.PerTypeScaleShift(
                len(type_names), per_type_energy_scales, per_type_energy_shifts
            )
            # pair potential after the scale/shift (allegro_models.py:270-288); only ZBL exists in nequip
            self.pair_potential = None
            if pair_potential is not None:
                pp = dict(pair_potential)
                target = pp.pop("_target_", "nequip.nn.pair_potential.ZBL").rsplit(".", 1)[-1]
                assert target == "ZBL", target
                self.pair_potential = R.ZBL(type_names=type_names, **pp)

    def forward(self, data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        data = self.edge_norm(data)
        data = self.radial_chemical_embed(data, self.model_dtype)
        data[R.EDGE_EMBEDDING_KEY] = self.scalar_embed_mlp(data[R.EDGE_EMBEDDING_KEY])
        data = self.tensor_embed(data)
        data = self.allegro(data)
        data[R.EDGE_ENERGY_KEY] = self.edge_readout(data[R.EDGE_FEATURES_KEY])
        data = self.edge_eng_sum(data)
        data = self.per_type_energy_scale_shift(data)
        if self.pair_potential is not None:
            data = self.pair_potential(data)
        data[R.TOTAL_ENERGY_KEY] = data[R.PER_ATOM_ENERGY_KEY].sum(dim=0, keepdim=True)
        return data


class AllegroOracle(torch.nn.Module):
    """AllegroModel = ForceStressOutput(AllegroEnergyModel) (allegro_models.py:101-103)."""

    def __init__(self, **kwargs):
        super().__init__()
        self.model = AllegroEnergyOracle(**kwargs)

    def forward(self, data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """nequip ForceStressOutput: forces = -dE/dpos; with a cell also stress = (dE/d eps)/V and
        virial = -dE/d eps, eps a symmetric infinitesimal strain applied to positions and cell."""
        data = dict(data)
        for k in (R.EDGE_VECTORS_KEY, R.EDGE_LENGTH_KEY):
            data.pop(k, None)
        pos = data[R.POSITIONS_KEY].detach().clone().requires_grad_(True)
        has_cell = R.CELL_KEY in data
        with torch.enable_grad():
            if has_cell:
                cell0 = data[R.CELL_KEY].view(3, 3).to(pos.dtype)
                disp = torch.zeros(3, 3, dtype=pos.dtype, device=pos.device, requires_grad=True)
                sym = 0.5 * (disp + disp.T)
                data[R.POSITIONS_KEY] = pos + pos @ sym
                data[R.CELL_KEY] = cell0 + cell0 @ sym
                data = self.model(data)
                g, gd = torch.autograd.grad(data[R.TOTAL_ENERGY_KEY].sum(), (pos, disp))
                volume = torch.linalg.det(cell0).abs()
                data[R.STRESS_KEY] = (gd / volume).unsqueeze(0)
                data[R.VIRIAL_KEY] = (-gd).unsqueeze(0)
                data[R.CELL_KEY] = cell0
            else:
                data[R.POSITIONS_KEY] = pos
                data = self.model(data)
                (g,) = torch.autograd.grad(data[R.TOTAL_ENERGY_KEY].sum(), pos)
        data[R.FORCE_KEY] = -g
        data[R.POSITIONS_KEY] = pos.detach()
        return {k: (v.detach() if isinstance(v, torch.Tensor) else v) for k, v in data.items()}
