"""CPU tests of the oracle modules: the reference's own test strategy restated without e3nn
(/root/reference/tests/nn/test_contract_basic.py, tests/nn/test_weighter.py,
tests/utils/test_compile_utils.py, tests/model/test_allegro.py:68-70)."""
import math

import numpy as np
import pytest
import torch

from allegro_b200 import data as D
from allegro_b200 import systems
from oracle import nn_ref as R
from oracle import o3_ref
from oracle.model_ref import AllegroOracle


def _block_D(irreps, Rm):
    """Block-diagonal representation matrix of O(3) element Rm (det may be -1) on ``irreps``."""
    det = float(torch.det(Rm))
    rot = Rm * det  # proper rotation
    blocks = []
    for _, ir in irreps:
        Dl = o3_ref.wigner_D_from_rotation(ir.l, rot)
        if det < 0:
            Dl = Dl * ir.p
        blocks.append(Dl)
    return torch.block_diag(*blocks)


def _brute_tp(c: R.Contracter, x1, x2):
    """Independent evaluation: loop over paths with explicit per-path CG blocks."""
    out = torch.zeros(x1.shape[0], c.mul, c.base_dim_out, dtype=x1.dtype)
    s1, s2, so = c.irreps_in1.slices(), c.irreps_in2.slices(), c.irreps_out.slices()
    instr = c.instructions
    if instr is None:
        instr = [(a, b, o) for o, (_, x) in enumerate(c.irreps_out) for a, (_, y) in enumerate(c.irreps_in1)
                 for b, (_, z) in enumerate(c.irreps_in2) if x in y * z]
    for p, (a, b, o) in enumerate(instr):
        l1, l2, l3 = c.irreps_in1[a][1].l, c.irreps_in2[b][1].l, c.irreps_out[o][1].l
        w = torch.from_numpy(np.array(o3_ref.wigner_3j(l1, l2, l3))).to(x1.dtype)
        if c.irrep_normalization == "component":
            w = w * math.sqrt(2 * l3 + 1)
        wt = c.weights
        if c.num_paths > 1:
            wt = wt[..., p]
        wt = wt.reshape(-1, 1) if c.path_channel_coupling else wt
        contrib = torch.einsum("ijk,zui,zuj->zuk", w, x1[:, :, s1[a]], x2[:, :, s2[b]])
        out[:, :, so[o]] += wt * contrib
    return out


@pytest.mark.parametrize("irreps_in1", ["0e + 0o + 1e + 1o", "2o + 1e + 0e", "0e+1o+2e"])
@pytest.mark.parametrize("irreps_in2", ["0e + 0o + 1e + 1o", "0e+1o+2e"])
@pytest.mark.parametrize("irreps_out", ["0e + 0o + 1e + 1o", "1o + 2e", "0e"])
@pytest.mark.parametrize("coupling", [True, False])
def test_contracter_matches_path_loop_and_is_equivariant(irreps_in1, irreps_in2, irreps_out, coupling):
    torch.manual_seed(0)
    torch.set_default_dtype(torch.float64)
    try:
        i1, i2, io = o3_ref.Irreps(irreps_in1), o3_ref.Irreps(irreps_in2), o3_ref.Irreps(irreps_out)
        try:
            c = R.Contracter(i1, i2, io, mul=3, path_channel_coupling=coupling, scatter_factor=0.37)
        except AssertionError:
            pytest.skip("no paths")
        E, N = 17, 5
        idx = torch.randint(0, N, (E,))
        x1, x2 = torch.randn(E, 3, i1.dim), torch.randn(E, 3, i2.dim)
        out = c(x1, x2, idx, N)
        x2g = 0.37 * R.scatter(x2, idx, N)[idx]
        assert (out - _brute_tp(c, x1, x2g)).abs().max() < 1e-12
        # equivariance under a random improper rotation
        Rm = -o3_ref.random_rotation(11)
        D1, D2, Do = _block_D(i1, Rm), _block_D(i2, Rm), _block_D(io, Rm)
        out_rot = c(x1 @ D1.T, x2 @ D2.T, idx, N)
        assert (out_rot - out @ Do.T).abs().max() < 1e-9
    finally:
        torch.set_default_dtype(torch.float32)


def test_contracter_gradcheck():
    torch.set_default_dtype(torch.float64)
    try:
        torch.manual_seed(1)
        i1 = o3_ref.Irreps("0e+1o+2e")
        c = R.Contracter(i1, i1, i1, mul=2, scatter_factor=0.5)
        idx = torch.randint(0, 3, (6,))
        x1 = torch.randn(6, 2, 9, requires_grad=True)
        x2 = torch.randn(6, 2, 9, requires_grad=True)
        assert torch.autograd.gradcheck(lambda a, b: c(a, b, idx, 3), (x1, x2), fast_mode=True)
    finally:
        torch.set_default_dtype(torch.float32)


@pytest.mark.parametrize("lmax", [1, 3])
@pytest.mark.parametrize("mul", [1, 5])
def test_weighter(lmax, mul):
    """MakeWeightedChannels == per-irrep scalar weighting (what e3nn Linear(shared_weights=False)
    does for 1 -> mul), and it commutes with rotations."""
    torch.manual_seed(2)
    irreps = o3_ref.Irreps.spherical_harmonics(lmax)
    m = R.MakeWeightedChannels(irreps, mul)
    x = torch.randn(7, irreps.dim, dtype=torch.float64)
    w = torch.randn(7, m.weight_numel, dtype=torch.float64)
    out = m(x, w)
    wv = w.view(7, mul, len(irreps))
    for r, sl in enumerate(irreps.slices()):
        assert (out[:, :, sl] - wv[:, :, r : r + 1] * x[:, None, sl]).abs().max() < 1e-14
    Dm = _block_D(irreps, o3_ref.random_rotation(5))
    assert (m(x @ Dm.T, w) - out @ Dm.T).abs().max() < 1e-12


def _model_and_data(name="c1", scale=None, dtype="float64", **over):
    d = systems.make_system(name, scale)
    kw = systems.model_kwargs(name, d[D.EDGE_INDEX_KEY].shape[1] / d[D.POSITIONS_KEY].shape[0], dtype)
    kw.update(over)
    return AllegroOracle(**kw), d


def test_model_invariance_and_force_equivariance():
    m, d = _model_and_data("c1")
    out = m(d)
    Rm = -o3_ref.random_rotation(3)
    d2 = dict(d)
    d2[D.POSITIONS_KEY] = d[D.POSITIONS_KEY] @ Rm.T
    d2[D.CELL_KEY] = d[D.CELL_KEY] @ Rm.T
    out2 = m(d2)
    assert (out[D.PER_ATOM_ENERGY_KEY] - out2[D.PER_ATOM_ENERGY_KEY]).abs().max() < 1e-10
    assert (out[D.FORCE_KEY] @ Rm.T - out2[D.FORCE_KEY]).abs().max() < 1e-10
    assert out[D.FORCE_KEY].sum(0).abs().max() < 1e-10  # translation invariance


def test_forces_are_minus_energy_gradient_fd():
    m, d = _model_and_data("c1")
    out = m(d)
    pos = d[D.POSITIONS_KEY]
    h = 1e-5
    for atom, ax in [(0, 0), (17, 2), (40, 1)]:
        dp, dm = dict(d), dict(d)
        pp, pm = pos.clone(), pos.clone()
        pp[atom, ax] += h
        pm[atom, ax] -= h
        dp[D.POSITIONS_KEY], dm[D.POSITIONS_KEY] = pp, pm
        fd = -(m(dp)[D.TOTAL_ENERGY_KEY] - m(dm)[D.TOTAL_ENERGY_KEY]).item() / (2 * h)
        assert fd == pytest.approx(out[D.FORCE_KEY][atom, ax].item(), abs=1e-7)


def test_strict_locality():
    """E_i depends only on atoms within r_max of i (tests/model/test_allegro.py:68-70)."""
    m, d = _model_and_data("c2", scale=3, l_max=2)
    pos = d[D.POSITIONS_KEY].clone().requires_grad_(True)
    dd = dict(d)
    dd[D.POSITIONS_KEY] = pos
    e = m.model(dd)[D.PER_ATOM_ENERGY_KEY]
    i = 5
    (g,) = torch.autograd.grad(e[i].sum(), pos)
    nb = set(d[D.EDGE_INDEX_KEY][1][d[D.EDGE_INDEX_KEY][0] == i].tolist()) | {i}
    far = [a for a in range(pos.shape[0]) if a not in nb]
    assert g[far].abs().max() == 0.0
    assert g[list(nb)].abs().max() > 0


def test_ghost_format_consistency():
    """PBC-with-shifts == ghost-atom format on owned atoms (allegro/_compile.py:17-65;
    tests/utils/test_compile_utils.py:7-18 checks the edge-length multiset)."""
    m, d = _model_and_data("c1")
    out = m(d)
    g = D.to_ghost_format(d)
    n = d[D.POSITIONS_KEY].shape[0]
    l0 = (d[D.POSITIONS_KEY][d[D.EDGE_INDEX_KEY][1]] - d[D.POSITIONS_KEY][d[D.EDGE_INDEX_KEY][0]]
          + d[D.EDGE_CELL_SHIFT_KEY] @ d[D.CELL_KEY]).norm(dim=-1)
    l1 = (g[D.POSITIONS_KEY][g[D.EDGE_INDEX_KEY][1]] - g[D.POSITIONS_KEY][g[D.EDGE_INDEX_KEY][0]]).norm(dim=-1)
    assert torch.allclose(l0.sort().values, l1.sort().values, atol=1e-12)
    g.pop("num_local_atoms")
    outg = m(g)
    assert (outg[D.PER_ATOM_ENERGY_KEY][:n] - out[D.PER_ATOM_ENERGY_KEY]).abs().max() < 1e-10


@pytest.mark.parametrize("cfg", [("c1", None), ("c2", 3), ("c5", 2)])
def test_model_runs_all_layer_shapes(cfg):
    name, scale = cfg
    over = dict(num_scalar_features=16, num_tensor_features=4, radial_chemical_embed_dim=16,
                scalar_embed_mlp_hidden_layers_width=16, allegro_mlp_hidden_layers_width=16,
                readout_mlp_hidden_layers_width=8)
    m, d = _model_and_data(name, scale, **over)
    out = m(d)
    assert torch.isfinite(out[D.FORCE_KEY]).all() and out[D.FORCE_KEY].abs().max() > 0


def test_stress_is_the_edge_virial_over_volume():
    """nequip ForceStressOutput's strain derivative (restated in AllegroOracle) equals the per-edge form the CUDA
    path uses: stress = sym(sum_z r_z (x) dE/dr_z) / V, virial = -sym(...)  (allegro_b200 energy_and_forces)."""
    import torch

    from allegro_b200 import systems
    from oracle import nn_ref as R
    from oracle.model_ref import AllegroOracle

    d = systems.make_system("c5", 2)
    kw = systems.model_kwargs("c5", 42.0, "float64")
    kw.update(num_scalar_features=8, num_tensor_features=4, radial_chemical_embed_dim=8, scalar_embed_mlp_hidden_layers_width=8,
              allegro_mlp_hidden_layers_width=8, readout_mlp_hidden_layers_width=8, l_max=2, num_layers=2,
              per_type_energy_scales=[1.0, 0.5, 2.0, 1.5, 0.25])
    oracle = AllegroOracle(**kw)
    out = oracle(d)
    pos, ei = d[R.POSITIONS_KEY], d[R.EDGE_INDEX_KEY]
    vec = (pos[ei[1]] - pos[ei[0]] + d[R.EDGE_CELL_SHIFT_KEY] @ d[R.CELL_KEY]).clone().requires_grad_(True)
    dd = {k: v for k, v in d.items()}
    dd[R.EDGE_VECTORS_KEY] = vec
    dd[R.EDGE_LENGTH_KEY] = vec.norm(dim=-1)
    with torch.enable_grad():
        e = oracle.model(dd)[R.TOTAL_ENERGY_KEY].sum()
        (g,) = torch.autograd.grad(e, vec)
    m = vec.detach().T @ g
    sym = 0.5 * (m + m.T)
    vol = torch.linalg.det(d[R.CELL_KEY]).abs()
    assert (sym / vol - out[R.STRESS_KEY][0]).abs().max() < 1e-12
    assert (-sym - out[R.VIRIAL_KEY][0]).abs().max() < 1e-10
    assert (m - m.T).abs().max() < 1e-10  # rotation invariance makes the raw edge virial symmetric already
