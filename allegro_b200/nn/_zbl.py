"""ZBL pair potential on top of the Allegro energy (reference call site allegro/model/allegro_models.py:270-288).

The module the reference instantiates there is nequip's ``nequip.nn.pair_potential.ZBL`` (LAMMPS ``pair_style zbl``,
constants of pair_zbl_const.h), preceded by an ``AddRadialCutoffToData(PolynomialCutoff(6))`` because no Allegro module
writes an edge cutoff.  State: one buffer ``atomic_numbers`` [num_types].  On the fused path the edge energies and their
position gradient come from one CUDA kernel (``ab2_zbl``, csrc/zbl.cu); the autograd path of ``forward`` uses the same
formula in torch ops on the device.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from .. import _lib
from ._modules import polynomial_cutoff

_SYMBOLS = (
    "X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr Rb Sr Y Zr Nb Mo Tc Ru Rh Pd "
    "Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn Fr Ra Ac Th "
    "Pa U Np Pu Am Cm Bk Cf Es Fm Md No Lr"
).split()
ATOMIC_NUMBERS = {s: z for z, s in enumerate(_SYMBOLS)}
QQR2E = {"metal": 14.399645, "real": 332.06371}  # LAMMPS force->qqr2e per unit system


class ZBL(torch.nn.Module):
    """kwargs of nequip's ZBL: ``type_names``, ``chemical_species`` (symbols per type; default: the type names), ``units``."""

    CUTOFF_P = 6.0  # PolynomialCutoff(6), allegro_models.py:275-277

    def __init__(self, type_names: Sequence[str], chemical_species: Optional[Sequence[str]] = None, units: str = "metal"):
        super().__init__()
        species = list(chemical_species) if chemical_species is not None else list(type_names)
        if len(species) != len(type_names):
            raise ValueError("ZBL: one chemical symbol per type")
        unknown = [s for s in species if s not in ATOMIC_NUMBERS]
        if unknown:
            raise ValueError(f"ZBL: unknown chemical symbols {unknown}")
        if units not in QQR2E:
            raise ValueError(f"ZBL: units must be one of {sorted(QQR2E)}")
        self.register_buffer("atomic_numbers", torch.tensor([float(ATOMIC_NUMBERS[s]) for s in species], dtype=torch.float64))
        self.units = units
        self.qq = QQR2E[units] * 0.5  # every pair is two directed edges

    # ---- fused path: one kernel, energies + gradient -------------------------------------
    def edge_energy_and_grad(self, vec, csr, types_i32, rmax_table, gvec: Optional[torch.Tensor]) -> torch.Tensor:
        Z = self.atomic_numbers.to(device=vec.device, dtype=vec.dtype)
        return _lib.zbl(self.CUTOFF_P, self.qq, vec, csr.ctr, csr.nbr, types_i32, Z, rmax_table.to(vec.dtype), gvec)

    # ---- autograd path --------------------------------------------------------------------
    def edge_energy(self, r: torch.Tensor, x_norm: torch.Tensor, type_c: torch.Tensor, type_n: torch.Tensor) -> torch.Tensor:
        Z = self.atomic_numbers.to(device=r.device, dtype=r.dtype)
        zi, zj = Z[type_c], Z[type_n]
        x = (zi.pow(0.23) + zj.pow(0.23)) * r / 0.46850
        psi = 0.02817 * torch.exp(-0.20162 * x) + 0.28022 * torch.exp(-0.40290 * x) + 0.50986 * torch.exp(-0.94229 * x) + 0.18175 * torch.exp(-3.19980 * x)
        return self.qq * zi * zj / r * psi * polynomial_cutoff(x_norm.to(r.dtype), self.CUTOFF_P)


def instantiate_pair_potential(spec, type_names) -> Optional[ZBL]:
    if spec is None:
        return None
    spec = dict(spec)
    target = str(spec.pop("_target_", "nequip.nn.pair_potential.ZBL")).rsplit(".", 1)[-1]
    if target != "ZBL":
        raise NotImplementedError(f"pair_potential {target!r}: only ZBL exists (nequip.nn.pair_potential)")
    spec.pop("type_names", None)
    spec.pop("irreps_in", None)
    return ZBL(type_names=type_names, **spec)
