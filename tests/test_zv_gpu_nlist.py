"""CUDA neighbour list straight into CSR (ab2_nl_bin / count / fill) against the torch cell list /
all-pairs search, and the prepared-frame route of the model (data["edge_csr"]) against the edge_index route."""
import pytest
import torch

from allegro_b200 import data as D
from allegro_b200 import systems

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _edge_set(ctr, nbr, shift_vec, cell):
    """canonical sorted rows (centre, neighbour, integer image) for set comparison"""
    inv = torch.linalg.inv(cell.double().cpu())
    img = torch.round(shift_vec.double().cpu() @ inv).long()
    rows = torch.cat([ctr.long().cpu().unsqueeze(1), nbr.long().cpu().unsqueeze(1), img], 1)
    order = torch.arange(rows.shape[0])
    for c in (4, 3, 2, 1, 0):
        order = order[torch.argsort(rows[order, c], stable=True)]
    return rows[order]


@pytest.mark.parametrize("pbc", [(True, True, True), (False, True, True), (True, False, False)])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_neighbor_csr_matches_torch_search(pbc, dtype):
    pos, cell, types = systems.make_positions("c2", 5)  # 18 A box, r_max 5
    g = torch.Generator().manual_seed(1)
    pos = pos + 0.3 * torch.randn(pos.shape, generator=g, dtype=pos.dtype)
    pos[7] += 2 * cell[1]          # raw MD coordinates: atoms outside the home cell
    pos[11] -= cell[2]
    if pbc[0]:
        pos[3] += cell[0]
    posd, celld = pos.to(DEV, dtype), cell.to(DEV, dtype)
    csr, sv = D.neighbor_csr(posd, 5.0, celld, pbc)
    ei, sh = D.neighbor_list(pos.to(dtype), 5.0, cell.to(dtype), pbc, method="brute")
    ref = _edge_set(ei[0], ei[1], sh.double() @ cell, cell)
    got = _edge_set(csr.ctr, csr.nbr, sv, cell)
    assert got.shape == ref.shape and torch.equal(got, ref)
    # rows are centre-sorted and row_ptr is consistent
    assert bool((csr.ctr[1:] >= csr.ctr[:-1]).all()) and int(csr.row_ptr[-1]) == csr.num_edges
    # edge vectors from the raw positions stay inside the cutoff
    v = posd[csr.nbr.long()] + sv - posd[csr.ctr.long()]
    assert float(v.norm(dim=-1).max()) < 5.0


def test_neighbor_csr_owned_centres_only():
    pos, cell, _ = systems.make_positions("c2", 5)
    posd = pos.to(DEV)
    csr_all, _ = D.neighbor_csr(posd, 5.0, cell.to(DEV))
    csr_own, sv = D.neighbor_csr(posd, 5.0, cell.to(DEV), n_centres=100)
    assert csr_own.num_atoms == 100 and csr_own.num_edges == int(csr_all.row_ptr[100])
    assert torch.equal(csr_own.nbr, csr_all.nbr[: csr_own.num_edges])


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-12), ("float32", 1e-4)])  # rows come in a different order: fp32 sums differ
def test_model_on_prepared_csr_equals_edge_index_route(dtype, tol):
    from test_gpu_model import _pair, _to_dev

    oracle, model, d = _pair("c2", 5, dtype)
    dd = _to_dev(d)
    ref = model(dd)
    csr, sv = D.neighbor_csr(dd[D.POSITIONS_KEY], 5.0, dd[D.CELL_KEY])
    prepared = {D.POSITIONS_KEY: dd[D.POSITIONS_KEY], D.ATOM_TYPE_KEY: dd[D.ATOM_TYPE_KEY], D.CELL_KEY: dd[D.CELL_KEY],
                D.CSR_KEY: csr, D.EDGE_SHIFT_VEC_KEY: sv}
    out = model(prepared)
    for k in (D.PER_ATOM_ENERGY_KEY, D.FORCE_KEY, D.STRESS_KEY):
        a, b = out[k].double(), ref[k].double()
        assert float((a - b).abs().max()) <= tol * float(b.abs().max()), k


def test_calculator_uses_the_cuda_list():
    from allegro_b200.calculator import AllegroCalculator
    from test_gpu_model import _pair, _to_dev

    oracle, model, d = _pair("c2", 5, "float64")
    dd = _to_dev(d)
    calc = AllegroCalculator(model, 5.0, skin=0.4, use_graph=True)
    out = calc.compute(dd[D.POSITIONS_KEY], dd[D.CELL_KEY], dd[D.ATOM_TYPE_KEY])
    assert D.CSR_KEY in calc._data and calc.num_edges > d[D.EDGE_INDEX_KEY].shape[1]  # skin list is larger
    ref = oracle(d)
    assert float((out["forces"].cpu() - ref[D.FORCE_KEY]).abs().max() / ref[D.FORCE_KEY].abs().max()) < 1e-9
    p2 = dd[D.POSITIONS_KEY] + 0.03 * torch.randn_like(dd[D.POSITIONS_KEY]).clamp(-3, 3)
    out2 = calc.compute(p2, dd[D.CELL_KEY])
    assert calc.n_rebuilds == 1  # inside the skin: the captured graph was replayed on the new positions
    d2 = dict(d)
    d2[D.POSITIONS_KEY] = p2.cpu()
    ei, sh = D.neighbor_list(d2[D.POSITIONS_KEY], 5.0, d[D.CELL_KEY])
    d2[D.EDGE_INDEX_KEY], d2[D.EDGE_CELL_SHIFT_KEY] = ei, sh
    ref2 = oracle(d2)
    assert float((out2["forces"].cpu() - ref2[D.FORCE_KEY]).abs().max() / ref2[D.FORCE_KEY].abs().max()) < 1e-9
