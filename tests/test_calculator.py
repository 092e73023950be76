"""Host logic of the MD-side driver (allegro_b200/calculator.py) with the CPU oracle standing in for the model:
skin-list evaluations equal exact-list evaluations, rebuilds happen exactly when an atom moved > skin/2."""
import torch

from allegro_b200 import data as D
from allegro_b200 import systems
from allegro_b200.calculator import AllegroCalculator
from oracle.model_ref import AllegroOracle

SMALL = dict(num_scalar_features=8, num_tensor_features=4, radial_chemical_embed_dim=8, scalar_embed_mlp_hidden_layers_width=8,
             allegro_mlp_hidden_layers_width=8, readout_mlp_hidden_layers_width=8)


def _setup(cfg="c5", scale=2, **over):
    d = systems.make_system(cfg, scale)
    kw = systems.model_kwargs(cfg, 40.0, "float64")
    kw.update(SMALL)
    kw.update(l_max=2, num_layers=2)
    kw.update(over)
    return AllegroOracle(**kw), d, systems.CONFIGS[cfg]["r_max"]


def _exact(oracle, pos, cell, types, r_max):
    ei, sh = D.neighbor_list(pos, r_max, cell, (True, True, True))
    return oracle({D.POSITIONS_KEY: pos, D.CELL_KEY: cell, D.ATOM_TYPE_KEY: types, D.EDGE_INDEX_KEY: ei, D.EDGE_CELL_SHIFT_KEY: sh})


def test_skin_edges_contribute_nothing_and_rebuild_rule():
    oracle, d, r_max = _setup()
    pos, cell, types = d[D.POSITIONS_KEY], d[D.CELL_KEY], d[D.ATOM_TYPE_KEY]
    calc = AllegroCalculator(oracle, r_max, skin=0.6, compute_stress=True)
    g = torch.Generator().manual_seed(11)
    p = pos.clone()
    rebuilds = []
    for step in range(6):
        out = calc.compute(p, cell, types)
        ref = _exact(oracle, p, cell, types, r_max)
        assert calc.num_edges > ref[D.EDGE_INDEX_KEY].shape[1]  # the skin list really is longer
        assert (out["energy"] - ref[D.TOTAL_ENERGY_KEY]).abs().max() < 1e-12
        assert (out["atomic_energy"] - ref[D.PER_ATOM_ENERGY_KEY]).abs().max() < 1e-12
        assert (out["forces"] - ref[D.FORCE_KEY]).abs().max() < 1e-12
        assert (out["stress"] - ref[D.STRESS_KEY]).abs().max() < 1e-12
        rebuilds.append(calc.n_rebuilds)
        p = p + 0.08 * torch.randn(p.shape, generator=g, dtype=p.dtype)  # random walk, ~0.14 A per step
    assert rebuilds[0] == 1 and rebuilds[1] == 1   # first step builds, small moves reuse the list
    assert rebuilds[-1] >= 2                        # ... until someone has moved more than skin/2 = 0.3 A
    assert calc.n_evaluations == 6


def test_rebuild_triggers():
    oracle, d, r_max = _setup("c2", 2)
    pos, cell, types = d[D.POSITIONS_KEY], d[D.CELL_KEY], d[D.ATOM_TYPE_KEY]
    calc = AllegroCalculator(oracle, r_max, skin=0.5)
    calc.compute(pos, cell, types)
    calc.compute(pos, cell)                                   # types may be omitted after the first call
    assert calc.n_rebuilds == 1
    p2 = pos.clone()
    p2[3, 0] += 0.2                                           # < skin/2
    calc.compute(p2, cell)
    assert calc.n_rebuilds == 1
    p2[3, 0] += 0.1                                           # 0.3 > skin/2
    calc.compute(p2, cell)
    assert calc.n_rebuilds == 2
    calc.compute(p2, cell * 1.01)                             # cell change always rebuilds
    assert calc.n_rebuilds == 3
    t2 = types.clone()
    calc.compute(p2, cell * 1.01, t2)                         # equal types: no rebuild
    assert calc.n_rebuilds == 3
    calc = AllegroCalculator(oracle, r_max, skin=0.5, check_every=3)
    calc.compute(pos, cell, types)
    p3 = pos.clone()
    p3[0, 1] += 0.4
    calc.compute(p3, cell)
    calc.compute(p3, cell)
    assert calc.n_rebuilds == 1                               # displacement only looked at every 3rd call
    calc.compute(p3, cell)
    assert calc.n_rebuilds == 2


def test_open_boundary_cluster():
    g = torch.Generator().manual_seed(2)
    pos = torch.rand(24, 3, generator=g, dtype=torch.float64) * 6.0
    types = torch.randint(0, 2, (24,), generator=g)
    kw = systems.model_kwargs("c2", 10.0, "float64")
    kw.update(SMALL)
    kw.update(type_names=["X", "Y"], r_max=3.5)
    oracle = AllegroOracle(**kw)
    calc = AllegroCalculator(oracle, 3.5, skin=0.4, pbc=False)
    out = calc.compute(pos, None, types)
    ei, _ = D.neighbor_list(pos, 3.5, None, (False, False, False))
    ref = oracle({D.POSITIONS_KEY: pos, D.ATOM_TYPE_KEY: types, D.EDGE_INDEX_KEY: ei})
    assert (out["forces"] - ref[D.FORCE_KEY]).abs().max() < 1e-12
    assert "stress" not in out
