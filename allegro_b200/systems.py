"""Seeded synthetic systems for BASELINE.json's configs.

c1  Si diamond 2x2x2 (64 atoms), r_max 4.0, l_max 1, L 1, S=W=32, U=16
c2  Cu FCC 14^3 (10 976 atoms), r_max 5.0, l_max 2, L 2, S=W=64, U=32   (bench default)
c3  Li-P-S-like 100 000 atoms (3 species), r_max 6.0, S=W=128, U=64
c4  water-like ~1M atoms, r_max 5.0, S=W=64, U=32 (8 slabs)
c5  5-species FCC 14^3, r_max 5.0, l_max 3, L 3, S=W=128, U=64, fp64

Positions are fp64, jittered U(-0.05, 0.05) A (seed 1234) so no edge sits on the cutoff
and forces are non-zero.  ``scale`` shrinks the supercell for CPU-sized parity cases.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from . import data as D

_FCC = torch.tensor([[0.0, 0.0, 0.0], [0.0, 0.5, 0.5], [0.5, 0.0, 0.5], [0.5, 0.5, 0.0]], dtype=torch.float64)
_DIAMOND = torch.cat([_FCC, _FCC + 0.25], 0)


def _lattice(basis: torch.Tensor, a: float, reps, jitter: float, gen: torch.Generator):
    nx, ny, nz = reps
    g = torch.stack(torch.meshgrid(torch.arange(nx), torch.arange(ny), torch.arange(nz), indexing="ij"), -1).reshape(-1, 1, 3).double()
    pos = ((g + basis.unsqueeze(0)) * a).reshape(-1, 3)
    pos = pos + (torch.rand(pos.shape, generator=gen, dtype=torch.float64) * 2 - 1) * jitter
    cell = torch.diag(torch.tensor([nx * a, ny * a, nz * a], dtype=torch.float64))
    return pos, cell


CONFIGS: Dict[str, Dict] = {
    "c1": dict(system="Si diamond 2x2x2", type_names=["Si"], r_max=4.0, l_max=1, num_layers=1, S=32, U=16, dtype="float32"),
    "c2": dict(system="Cu FCC 14^3", type_names=["Cu"], r_max=5.0, l_max=2, num_layers=2, S=64, U=32, dtype="bfloat16"),
    "c3": dict(system="LiPS-like 100k", type_names=["Li", "P", "S"], r_max=6.0, l_max=2, num_layers=2, S=128, U=64, dtype="float32"),
    "c4": dict(system="water-like 1M", type_names=["O", "H"], r_max=5.0, l_max=2, num_layers=2, S=64, U=32, dtype="float32"),
    "c5": dict(system="HEA FCC 14^3", type_names=["A", "B", "C", "D", "E"], r_max=5.0, l_max=3, num_layers=3, S=128, U=64, dtype="float64"),
}


def make_positions(name: str, scale=None, seed: int = 1234):
    """-> pos [N,3] fp64, cell [3,3] fp64, atom_types [N] int64."""
    gen = torch.Generator().manual_seed(seed)

    def _reps(default):
        if scale is None:
            return (default,) * 3
        return tuple(scale) if isinstance(scale, (tuple, list)) else (scale,) * 3

    if name == "c1":
        pos, cell = _lattice(_DIAMOND, 5.431, _reps(2), 0.05, gen)
        types = torch.zeros(pos.shape[0], dtype=torch.long)
    elif name == "c2":
        pos, cell = _lattice(_FCC, 3.615, _reps(14), 0.05, gen)
        types = torch.zeros(pos.shape[0], dtype=torch.long)
    elif name == "c5":
        pos, cell = _lattice(_FCC, 3.6, _reps(14), 0.05, gen)
        types = torch.randint(0, 5, (pos.shape[0],), generator=gen)
    elif name == "c3":
        # jittered simple-cubic lattice at number density 0.050 A^-3 (spacing 2.714 A);
        # jitter 0.3 A keeps min distance > 1.8 A.  Li:P:S = 3:1:4.
        a = (1.0 / 0.050) ** (1.0 / 3.0)  # 46^3 = 97 336 ~ 100k
        pos, cell = _lattice(torch.zeros(1, 3, dtype=torch.float64), a, _reps(46), 0.3, gen)
        r = torch.rand(pos.shape[0], generator=gen)
        types = torch.where(r < 3 / 8, 0, torch.where(r < 4 / 8, 1, 2)).long()
    elif name == "c4":
        # O on a jittered cubic lattice at 0.0334 A^-3 + 2 H at 0.96 A in random directions
        a = (1.0 / 0.0334) ** (1.0 / 3.0)  # 69^3 = 328 509 O -> 985 527 atoms
        o, cell = _lattice(torch.zeros(1, 3, dtype=torch.float64), a, _reps(69), 0.2, gen)
        d1 = torch.randn(o.shape, generator=gen, dtype=torch.float64)
        d1 = d1 / d1.norm(dim=-1, keepdim=True)
        d2 = torch.randn(o.shape, generator=gen, dtype=torch.float64)
        d2 = d2 - (d2 * d1).sum(-1, keepdim=True) * d1
        d2 = d2 / d2.norm(dim=-1, keepdim=True)
        ang = math.radians(104.5)
        h1 = o + 0.96 * d1
        h2 = o + 0.96 * (math.cos(ang) * d1 + math.sin(ang) * d2)
        pos = torch.stack([o, h1, h2], 1).reshape(-1, 3)
        types = torch.tensor([0, 1, 1]).repeat(o.shape[0])
    else:
        raise KeyError(name)
    return pos, cell, types


def make_system(name: str, scale=None, seed: int = 1234, device="cpu") -> D.Type:
    cfg = CONFIGS[name]
    pos, cell, types = make_positions(name, scale, seed)
    pos, cell, types = pos.to(device), cell.to(device), types.to(device)
    ei, shift = D.neighbor_list(pos, cfg["r_max"], cell, (True, True, True))
    return {
        D.POSITIONS_KEY: pos,
        D.CELL_KEY: cell,
        D.ATOM_TYPE_KEY: types,
        D.EDGE_INDEX_KEY: ei,
        D.EDGE_CELL_SHIFT_KEY: shift,
    }


def model_kwargs(name: str, avg_num_neighbors: float, model_dtype: Optional[str] = None, seed: int = 456) -> Dict:
    """AllegroModel kwargs of a config (reference kwarg names, allegro_models.py:112-151)."""
    c = CONFIGS[name]
    return dict(
        seed=seed,
        model_dtype=model_dtype or c["dtype"],
        type_names=c["type_names"],
        r_max=c["r_max"],
        l_max=c["l_max"],
        parity=True,
        radial_chemical_embed={"_target_": "allegro.nn.TwoBodyBesselScalarEmbed", "num_bessels": 8, "polynomial_cutoff_p": 6},
        radial_chemical_embed_dim=c["S"],
        scalar_embed_mlp_hidden_layers_depth=1,
        scalar_embed_mlp_hidden_layers_width=c["S"],
        num_layers=c["num_layers"],
        num_scalar_features=c["S"],
        num_tensor_features=c["U"],
        allegro_mlp_hidden_layers_depth=1,
        allegro_mlp_hidden_layers_width=c["S"],
        tp_path_channel_coupling=True,
        readout_mlp_hidden_layers_depth=1,
        readout_mlp_hidden_layers_width=c["S"],
        avg_num_neighbors=avg_num_neighbors,
    )
