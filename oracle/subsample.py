"""Oracle on a SUB-SAMPLE of a large frame (test infrastructure; see oracle/__init__.py).

Allegro is strictly local (/root/reference/tests/model/test_allegro.py:68-70; every layer's
environment sum runs over the edges of ONE centre, allegro/nn/_strided/_contract.py:199-205):
  E_i depends only on the edges centred on i;
  F_i = -dE/dpos_i depends only on the edges centred on i and on the neighbours of i.
So the reference values of a set A of atoms of an arbitrarily large frame come from the oracle
evaluated on the edge subset  {z : centre(z) in A u N(A)}  with the FULL position array: E_i is exact
for every centre kept, F_i for every i in A.  This is how the named full-size configurations
(10^4..10^6 atoms) are checked without running the CPU oracle on 10^7 edges.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

_POS, _EI, _SHIFT, _E, _F, _EA = "pos", "edge_index", "edge_cell_shift", "total_energy", "forces", "atomic_energy"


def local_reference(oracle, data: Dict[str, torch.Tensor], atoms: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (centres kept [C], their atomic energies [C,1], forces of ``atoms`` [A,3]); CPU fp64 oracle."""
    d = {k: (v.detach().cpu() if torch.is_tensor(v) else v) for k, v in data.items()}
    ei = d[_EI]
    n = d[_POS].shape[0]
    atoms = atoms.cpu().long()
    in_a = torch.zeros(n, dtype=torch.bool)
    in_a[atoms] = True
    keep_c = in_a.clone()
    keep_c[ei[1][in_a[ei[0]]]] = True  # N(A): neighbours of the atoms of A (the list is symmetric)
    sel = keep_c[ei[0]]
    sub = dict(d)
    sub[_EI] = ei[:, sel].contiguous()
    if _SHIFT in d:
        sub[_SHIFT] = d[_SHIFT][sel].contiguous()
    out = oracle(sub)
    centres = keep_c.nonzero().reshape(-1)
    return centres, out[_EA][centres], out[_F][atoms]


def ball(pos: torch.Tensor, n_atoms: int, seed: int = 0) -> torch.Tensor:
    """Indices of the ``n_atoms`` atoms closest to a seeded random atom (a compact sample keeps A u N(A) small)."""
    g = torch.Generator().manual_seed(seed)
    p = pos.detach().cpu().double()
    c = p[int(torch.randint(0, p.shape[0], (1,), generator=g))]
    return torch.argsort((p - c).norm(dim=-1))[:n_atoms]
