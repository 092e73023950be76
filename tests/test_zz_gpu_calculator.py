"""MD-side driver on the GPU: Verlet-skin list + CUDA-graph replay (allegro_b200/calculator.py, graph.py) against
the oracle evaluated on exact r_max neighbour lists along a short random walk."""
import pytest
import torch

from allegro_b200 import data as D

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _exact(oracle, pos, cell, types, r_max):
    ei, sh = D.neighbor_list(pos, r_max, cell, (True, True, True))
    return oracle({D.POSITIONS_KEY: pos, D.CELL_KEY: cell, D.ATOM_TYPE_KEY: types, D.EDGE_INDEX_KEY: ei, D.EDGE_CELL_SHIFT_KEY: sh})


@pytest.mark.parametrize("use_graph", [False, True], ids=["eager", "graph"])
def test_calculator_random_walk(use_graph):
    from allegro_b200.calculator import AllegroCalculator
    from test_gpu_model import _pair

    oracle, model, d = _pair("c2", 3, "float64")
    pos, cell, types = d[D.POSITIONS_KEY], d[D.CELL_KEY], d[D.ATOM_TYPE_KEY]
    calc = AllegroCalculator(model, 5.0, skin=0.6, use_graph=use_graph, compute_stress=not use_graph)
    assert calc.use_graph == use_graph
    g = torch.Generator().manual_seed(4)
    p = pos.clone()
    for step in range(5):
        out = calc.compute(p.to(DEV), cell.to(DEV), types.to(DEV))
        ref = _exact(oracle, p, cell, types, 5.0)
        f, e = out["forces"].double().cpu(), out["atomic_energy"].double().cpu()
        assert (f - ref[D.FORCE_KEY]).abs().max() / ref[D.FORCE_KEY].abs().max() < 1e-9
        assert (e - ref[D.PER_ATOM_ENERGY_KEY]).abs().max() / ref[D.PER_ATOM_ENERGY_KEY].abs().max() < 1e-9
        if not use_graph:
            assert (out["stress"].double().cpu() - ref[D.STRESS_KEY]).abs().max() / ref[D.STRESS_KEY].abs().max() < 1e-9
        p = p + 0.1 * torch.randn(p.shape, generator=g, dtype=p.dtype)
    assert calc.n_evaluations == 5 and 2 <= calc.n_rebuilds <= 5
