"""AtomicDataDict keys, neighbour lists and the centre-sorted CSR edge format.

Key strings are nequip's ``AtomicDataDict`` constants (the three output
keys are confirmed by /root/reference/tests/model/test_allegro.py:233).

The on-device edge format every kernel consumes is *centre-sorted CSR*: edges ordered by
``edge_index[0]`` (the centre, allegro/nn/_allegro.py:238), ``row_ptr[N+1]`` int32,
``ctr[E]``/``nbr[E]`` int32.  With it, the per-centre environment sum of
allegro/nn/_strided/_contract.py:199-205 is a reduction over a contiguous edge range.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

POSITIONS_KEY = "pos"
EDGE_INDEX_KEY = "edge_index"
ATOM_TYPE_KEY = "atom_types"
CELL_KEY = "cell"
PBC_KEY = "pbc"
EDGE_CELL_SHIFT_KEY = "edge_cell_shift"
BATCH_KEY = "batch"
NUM_NODES_KEY = "num_atoms"
EDGE_VECTORS_KEY = "edge_vectors"
EDGE_LENGTH_KEY = "edge_lengths"
NORM_LENGTH_KEY = "normed_edge_lengths"
EDGE_TYPE_KEY = "edge_type"
EDGE_ATTRS_KEY = "edge_attrs"
EDGE_EMBEDDING_KEY = "edge_embedding"
EDGE_FEATURES_KEY = "edge_features"
EDGE_CUTOFF_KEY = "edge_cutoff"
EDGE_ENERGY_KEY = "edge_energy"
PER_ATOM_ENERGY_KEY = "atomic_energy"
TOTAL_ENERGY_KEY = "total_energy"
FORCE_KEY = "forces"
STRESS_KEY = "stress"
VIRIAL_KEY = "virial"
# prepared-frame keys (this package only): a prebuilt centre-sorted CSR (EdgeCSR) and the per-edge shift vectors
# [E,3] = edge_cell_shift @ cell in CSR order.  When present they are used instead of edge_index / edge_cell_shift,
# so the int64 COO list never has to exist (neighbor_csr below; 5x10^7 edges at the 1M-atom scale).
CSR_KEY = "edge_csr"
EDGE_SHIFT_VEC_KEY = "edge_shift_vec"

Type = Dict[str, torch.Tensor]


def num_nodes(data: Type) -> int:
    return data[POSITIONS_KEY].shape[0]


# --------------------------------------------------------------------------- #
# neighbour lists
# --------------------------------------------------------------------------- #
def _brute_force(pos, cell, pbc, r_max):
    """All-pairs search over the periodic images within r_max.  Positions are first wrapped into the
    cell along the periodic axes (MD drivers hand over unwrapped coordinates; without the wrap, atoms
    that diffused by a lattice vector would lose neighbours) and the integer image offsets are folded
    back into the returned shifts, so  r = pos[j] + shift @ cell - pos[i]  holds for the RAW
    positions -- the same convention as ``_cell_list``."""
    dev = pos.device
    if cell is None or not any(pbc):
        shifts = torch.zeros(1, 3, dtype=torch.long, device=dev)
        cell_m = torch.zeros(3, 3, dtype=pos.dtype, device=dev)
        img0 = torch.zeros(pos.shape[0], 3, dtype=torch.long, device=dev)
        wrapped = pos
    else:
        cell_m = cell.view(3, 3).to(pos.dtype)
        pbc_t = torch.tensor([bool(p) for p in pbc], device=dev)
        frac = torch.linalg.solve(cell_m.T, pos.T).T  # pos = frac @ cell
        img0 = torch.where(pbc_t, torch.floor(frac), torch.zeros_like(frac)).to(torch.long)
        wrapped = pos - img0.to(pos.dtype) @ cell_m
        # number of images needed per axis: r_max / (height of the cell along that axis)
        vol = torch.det(cell_m).abs()
        cr = torch.stack([torch.linalg.cross(cell_m[1], cell_m[2]), torch.linalg.cross(cell_m[2], cell_m[0]), torch.linalg.cross(cell_m[0], cell_m[1])])
        heights = vol / cr.norm(dim=-1)
        reps = [int(math.ceil(r_max / float(h))) if p else 0 for h, p in zip(heights, pbc)]
        rng = [torch.arange(-r, r + 1, device=dev) for r in reps]
        shifts = torch.stack(torch.meshgrid(*rng, indexing="ij"), dim=-1).reshape(-1, 3)
    ei, sh = [], []
    for s in shifts:
        off = s.to(pos.dtype) @ cell_m
        d = wrapped.unsqueeze(0) + off - wrapped.unsqueeze(1)  # [i, j]
        mask = d.norm(dim=-1) < r_max
        if not bool(s.any()):
            mask.fill_diagonal_(False)
        ij = mask.nonzero()
        ei.append(ij.T)
        sh.append(s.expand(ij.shape[0], 3) - img0[ij[:, 1]] + img0[ij[:, 0]])
    return torch.cat(ei, dim=1), torch.cat(sh, dim=0)


def _cell_list(pos, box, r_max, pbc=(True, True, True), origin=None):
    """Orthorhombic box, per-axis periodicity.  Periodic axes need >= 3 cells; on a
    non-periodic axis the grid spans [origin, origin+box) and nothing wraps."""
    dev = pos.device
    n = pos.shape[0]
    pbc_t = torch.tensor([bool(p) for p in pbc], device=dev)
    if origin is None:
        origin = torch.zeros(3, dtype=pos.dtype, device=dev)
    ncell = torch.floor(box / r_max).to(torch.long).clamp(min=1)
    assert int(ncell[pbc_t].min() if bool(pbc_t.any()) else 3) >= 3, "cell list needs box >= 3 r_max on every periodic axis"
    rel = pos - origin
    img0 = torch.where(pbc_t, torch.floor(rel / box), torch.zeros_like(rel)).to(torch.long)  # image index of the raw position
    wrapped = rel - img0.to(pos.dtype) * box
    cidx3 = torch.minimum(torch.clamp((wrapped / box * ncell).to(torch.long), min=0), ncell - 1)
    nc = [int(v) for v in ncell]
    cid = (cidx3[:, 0] * nc[1] + cidx3[:, 1]) * nc[2] + cidx3[:, 2]
    order = torch.argsort(cid, stable=True)
    ncells = nc[0] * nc[1] * nc[2]
    counts = torch.bincount(cid, minlength=ncells)
    starts = torch.cumsum(counts, 0) - counts
    ei, sh = [], []
    ar = torch.arange(n, device=dev)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                d = torch.tensor([dx, dy, dz], device=dev)
                c3 = cidx3 + d
                img = torch.div(c3, ncell, rounding_mode="floor")  # -1, 0, +1
                valid = (pbc_t | (img == 0)).all(-1)
                img = torch.where(pbc_t, img, torch.zeros_like(img))
                c3w = torch.clamp(c3 - img * ncell, min=0)
                c3w = torch.minimum(c3w, ncell - 1)
                ncid = (c3w[:, 0] * nc[1] + c3w[:, 1]) * nc[2] + c3w[:, 2]
                cnt = counts[ncid] * valid
                tot = int(cnt.sum())
                if tot == 0:
                    continue
                i_rep = torch.repeat_interleave(ar, cnt)
                offs = torch.arange(tot, device=dev) - torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt)
                j_rep = order[starts[ncid][i_rep] + offs]
                img_rep = img[i_rep]
                rij = wrapped[j_rep] + img_rep.to(pos.dtype) * box - wrapped[i_rep]
                keep = (rij.norm(dim=-1) < r_max) & ~((i_rep == j_rep) & (img_rep == 0).all(-1))
                i_k, j_k = i_rep[keep], j_rep[keep]
                # shift relative to the *raw* positions: r = pos[j] + shift*box - pos[i]
                s = img_rep[keep] - img0[j_k] + img0[i_k]
                ei.append(torch.stack([i_k, j_k]))
                sh.append(s)
    return torch.cat(ei, dim=1), torch.cat(sh, dim=0)


def neighbor_list(
    pos: torch.Tensor,
    r_max: float,
    cell: Optional[torch.Tensor] = None,
    pbc=(True, True, True),
    method: str = "auto",
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Full (directed) neighbour list, sorted by centre then neighbour.
    Returns edge_index [2,E] int64 (row 0 = centre) and edge_cell_shift [E,3] (pos dtype)."""
    pbc = tuple(bool(p) for p in (pbc if not isinstance(pbc, bool) else (pbc,) * 3))
    ortho = cell is not None and bool((cell.view(3, 3) - torch.diag(torch.diagonal(cell.view(3, 3)))).abs().max() == 0)
    if method == "auto":
        big = pos.shape[0] > 3000
        can = cell is not None and ortho and all(
            (not p) or float(cell.view(3, 3)[a, a]) >= 3 * r_max for a, p in enumerate(pbc)
        )
        method = "cell" if (big and can) else "brute"
    if method == "cell":
        box = torch.diagonal(cell.view(3, 3)).to(pos.dtype).clone()
        origin = torch.zeros(3, dtype=pos.dtype, device=pos.device)
        for a, p in enumerate(pbc):
            if not p:  # non-periodic axis: grid over the occupied extent
                lo, hi = pos[:, a].min(), pos[:, a].max()
                origin[a] = lo
                box[a] = (hi - lo) * (1 + 1e-9) + 1e-6
        ei, sh = _cell_list(pos, box, float(r_max), pbc, origin)
    else:
        ei, sh = _brute_force(pos, cell, pbc, float(r_max))
    n = pos.shape[0]
    # sort by (centre, neighbour); ties (same pair through different images) keep a deterministic order
    # via the shift.  Successive stable sorts from the least significant key (any shift magnitude --
    # raw MD positions may sit many cells away from the home cell).
    order = torch.arange(ei.shape[1], device=ei.device)
    for k in (sh[:, 2], sh[:, 1], sh[:, 0], ei[0] * n + ei[1]):
        order = order[torch.argsort(k[order], stable=True)]
    return ei[:, order].contiguous(), sh[order].to(pos.dtype).contiguous()


# --------------------------------------------------------------------------- #
# CSR edge format
# --------------------------------------------------------------------------- #
class EdgeCSR:
    """Centre-sorted edge list. ``perm`` maps sorted position -> original edge (None if the
    input was already sorted)."""

    __slots__ = ("num_atoms", "num_edges", "ctr", "nbr", "row_ptr", "perm", "max_degree", "_transposed")

    def __init__(self, num_atoms, ctr, nbr, row_ptr, perm, max_degree):
        self.num_atoms = int(num_atoms)
        self.num_edges = int(ctr.shape[0])
        self.ctr, self.nbr, self.row_ptr, self.perm = ctr, nbr, row_ptr, perm
        self.max_degree = int(max_degree)
        self._transposed = None

    def transposed(self, n_total: int):
        """Transposed CSR for the neighbour-side force reduction (ab2_force_scatter): ``col_ptr``
        [n_total+1] int32 and ``col_perm`` [E] int32 = edge ids grouped by neighbour atom, in
        ascending edge order inside a group (stable sort -> a fixed summation order).  Built once
        per neighbour list; ``n_total`` counts owned + ghost atoms."""
        if self._transposed is None or self._transposed[0] != int(n_total):
            nbr64 = self.nbr.long()
            col_perm = torch.argsort(nbr64, stable=True).to(torch.int32).contiguous()
            counts = torch.bincount(nbr64, minlength=int(n_total))
            if counts.shape[0] != int(n_total):
                raise ValueError("edge neighbour index out of range")
            col_ptr = torch.zeros(int(n_total) + 1, dtype=torch.int32, device=self.nbr.device)
            col_ptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
            self._transposed = (int(n_total), col_ptr, col_perm)
        return self._transposed[1], self._transposed[2]


def build_csr(edge_index: torch.Tensor, num_centres: int) -> EdgeCSR:
    """Sort edges by centre (stable) and build row_ptr.  ``num_centres`` = number of atoms that
    may be centres (owned atoms); neighbour indices may exceed it (ghost atoms,
    /root/reference/allegro/_compile.py:41-61)."""
    ctr64, nbr64 = edge_index[0], edge_index[1]
    E = ctr64.shape[0]
    if E > 0 and bool((ctr64[1:] >= ctr64[:-1]).all()):
        perm = None
    else:
        perm = torch.argsort(ctr64, stable=True)
        ctr64, nbr64 = ctr64[perm], nbr64[perm]
    counts = torch.bincount(ctr64, minlength=num_centres)
    if counts.shape[0] != num_centres:
        raise ValueError("edge centre index out of range")
    row_ptr = torch.zeros(num_centres + 1, dtype=torch.int32, device=edge_index.device)
    row_ptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
    maxdeg = int(counts.max()) if E > 0 else 0
    return EdgeCSR(num_centres, ctr64.to(torch.int32).contiguous(), nbr64.to(torch.int32).contiguous(), row_ptr, perm, maxdeg)


def to_ghost_format(data: Type) -> Type:
    """PBC (cell shifts) -> appended ghost atoms, the pair_allegro data contract
    (/root/reference/allegro/_compile.py:17-65): ghost pos = pos[j] + shift @ cell, ghost index
    = N + arange, inside-cell edges first.  Also the single-r_max halo format used for the
    multi-GPU decomposition."""
    data = dict(data)
    data.pop(BATCH_KEY, None)
    data.pop(NUM_NODES_KEY, None)
    if EDGE_CELL_SHIFT_KEY not in data:
        return data
    pos, ei = data[POSITIONS_KEY], data[EDGE_INDEX_KEY]
    shift, cell = data[EDGE_CELL_SHIFT_KEY], data[CELL_KEY].view(3, 3)
    outside = shift.abs().sum(-1) != 0
    ei_out = ei[:, outside].clone()
    pos_out = pos[ei_out[1]] + shift[outside].to(pos.dtype) @ cell.to(pos.dtype)
    typ = data[ATOM_TYPE_KEY].reshape(-1)
    typ_out = typ[ei_out[1]]
    ei_out[1] = torch.arange(pos.shape[0], pos.shape[0] + pos_out.shape[0], device=pos.device)
    data[POSITIONS_KEY] = torch.cat([pos, pos_out], 0)
    data[ATOM_TYPE_KEY] = torch.cat([typ, typ_out], 0)
    data[EDGE_INDEX_KEY] = torch.cat([ei[:, ~outside], ei_out], 1)
    data.pop(EDGE_CELL_SHIFT_KEY)
    data.pop(CELL_KEY)
    data.pop(PBC_KEY, None)
    data["num_local_atoms"] = torch.tensor(pos.shape[0])
    return data


def csr_supported(pos: torch.Tensor, r_max: float, cell: Optional[torch.Tensor], pbc=(True, True, True)) -> bool:
    """Can ``neighbor_csr`` (CUDA cell list) take this frame?  CUDA positions, orthorhombic box, >= 3 cells of
    edge r_max on every periodic axis."""
    if not pos.is_cuda or cell is None:
        return False
    c = cell.view(3, 3)
    if bool((c - torch.diag(torch.diagonal(c))).abs().max() != 0):
        return False
    return all((not p) or float(c[a, a]) >= 3 * r_max for a, p in enumerate(pbc))


def neighbor_csr(pos: torch.Tensor, r_max: float, cell: torch.Tensor, pbc=(True, True, True), n_centres: Optional[int] = None):
    """Neighbour search on the device straight into the kernels' format (ab2_nl_bin / count / fill).
    -> (EdgeCSR, shift_vec [E,3] in the positions' dtype).  Centres are atoms [0, n_centres) (owned atoms first)."""
    from . import _lib

    pbc = tuple(bool(p) for p in (pbc if not isinstance(pbc, bool) else (pbc,) * 3))
    if not csr_supported(pos, r_max, cell, pbc):
        raise ValueError("neighbor_csr needs CUDA positions and an orthorhombic box with >= 3 r_max per periodic axis")
    box = [float(v) for v in torch.diagonal(cell.view(3, 3))]
    origin = [0.0, 0.0, 0.0]
    for a, p in enumerate(pbc):
        if not p:  # open axis: grid over the occupied extent
            lo, hi = float(pos[:, a].min()), float(pos[:, a].max())
            origin[a] = lo
            box[a] = (hi - lo) * (1 + 1e-9) + 1e-6
    n = pos.shape[0]
    nc = n if n_centres is None else int(n_centres)
    row_ptr, nbr, shift = _lib.neighbor_csr(pos, r_max, box, pbc, origin, nc)
    counts = (row_ptr[1:] - row_ptr[:-1])
    ctr = torch.repeat_interleave(torch.arange(nc, device=pos.device, dtype=torch.int32), counts.long())
    maxdeg = int(counts.max()) if nc > 0 else 0
    return EdgeCSR(nc, ctr.contiguous(), nbr, row_ptr, None, maxdeg), shift
