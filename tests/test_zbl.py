"""ZBL pair potential (reference call site allegro/model/allegro_models.py:270-288).

nequip's module is not vendored, so the oracle restates the published algorithm (LAMMPS pair_style zbl) -- parity unpinned.
What can be pinned here: the oracle against an independent closed-form evaluation on a dimer (plain ``math``), its forces
against finite differences, and the product (host logic on CPU through the executable kernel spec; the CUDA kernel on the
GPU) against the oracle."""
import math

import pytest
import torch

import kernel_spec
from allegro_b200 import data as D
from allegro_b200 import systems
from oracle.model_ref import AllegroOracle

PAIR = {"_target_": "nequip.nn.pair_potential.ZBL", "units": "metal", "chemical_species": ["Li", "P", "S"]}


def _kw(dtype, avg):
    kw = systems.model_kwargs("c3", avg, dtype)
    kw.update(num_scalar_features=16, num_tensor_features=8, radial_chemical_embed_dim=16, scalar_embed_mlp_hidden_layers_width=16,
              allegro_mlp_hidden_layers_width=16, readout_mlp_hidden_layers_width=16, pair_potential=dict(PAIR))
    return kw


def _frame():
    d = systems.make_system("c3", 3)
    return d, d[D.EDGE_INDEX_KEY].shape[1] / d[D.POSITIONS_KEY].shape[0]


def test_oracle_zbl_dimer_closed_form():
    """Two Cu atoms 1.3 A apart in a big box: total ZBL energy = qqr2e Z^2 / r phi(x) u(r / r_max)."""
    r, rmax, Zc = 1.3, 4.0, 29.0
    x = 2 * Zc**0.23 * r / 0.46850
    phi = 0.18175 * math.exp(-3.19980 * x) + 0.50986 * math.exp(-0.94229 * x) + 0.28022 * math.exp(-0.40290 * x) + 0.02817 * math.exp(-0.20162 * x)
    t = r / rmax
    u = 1 - 28 * t**6 + 48 * t**7 - 21 * t**8
    want = 14.399645 * Zc * Zc / r * phi * u
    kw = systems.model_kwargs("c2", 1.0, "float64")
    kw.update(r_max=rmax, pair_potential={"_target_": "nequip.nn.pair_potential.ZBL", "units": "metal", "chemical_species": ["Cu"]},
              per_type_energy_scales=[0.0])  # network contribution scaled to zero: what is left is the pair term
    oracle = AllegroOracle(**kw)
    pos = torch.tensor([[0.0, 0.0, 0.0], [r, 0.0, 0.0]], dtype=torch.float64)
    data = {D.POSITIONS_KEY: pos, D.ATOM_TYPE_KEY: torch.zeros(2, dtype=torch.long), D.EDGE_INDEX_KEY: torch.tensor([[0, 1], [1, 0]])}
    out = oracle(data)
    assert abs(float(out[D.TOTAL_ENERGY_KEY]) - want) < 1e-10 * abs(want)
    # force on atom 0 along -x equals -dE/dr by central differences of the closed form
    def e_of(rr):
        xx = 2 * Zc**0.23 * rr / 0.46850
        ph = 0.18175 * math.exp(-3.19980 * xx) + 0.50986 * math.exp(-0.94229 * xx) + 0.28022 * math.exp(-0.40290 * xx) + 0.02817 * math.exp(-0.20162 * xx)
        tt = rr / rmax
        return 14.399645 * Zc * Zc / rr * ph * (1 - 28 * tt**6 + 48 * tt**7 - 21 * tt**8)

    dedr = (e_of(r + 1e-5) - e_of(r - 1e-5)) / 2e-5
    assert abs(float(out[D.FORCE_KEY][0, 0]) - dedr) < 1e-6 * abs(dedr)
    assert abs(float(out[D.FORCE_KEY][1, 0]) + dedr) < 1e-6 * abs(dedr)


def test_kernel_spec_zbl_matches_oracle_module():
    from oracle import nn_ref as R

    d, _ = _frame()
    ei = d[D.EDGE_INDEX_KEY]
    vec = d[D.POSITIONS_KEY][ei[1]] - d[D.POSITIONS_KEY][ei[0]] + d[D.EDGE_CELL_SHIFT_KEY].double() @ d[D.CELL_KEY]
    types = d[D.ATOM_TYPE_KEY]
    z = R.ZBL(["Li", "P", "S"], PAIR["chemical_species"])
    dd = {R.EDGE_LENGTH_KEY: vec.norm(dim=-1), R.EDGE_INDEX_KEY: ei, R.EDGE_TYPE_KEY: types[ei], R.NORM_LENGTH_KEY: (vec.norm(dim=-1) / 6.0).unsqueeze(-1),
          R.PER_ATOM_ENERGY_KEY: torch.zeros(types.shape[0], 1, dtype=torch.float64)}
    want = z(dd)[R.PER_ATOM_ENERGY_KEY].squeeze(-1)
    rmax = torch.full((3, 3), 6.0, dtype=torch.float64)
    ez = kernel_spec.zbl(6.0, 14.399645 * 0.5, vec, ei[0], ei[1], types, z.atomic_numbers, rmax, None)
    got = torch.zeros_like(want).index_add_(0, ei[0], ez)
    assert float((got - want).abs().max()) < 1e-12 * float(want.abs().max())


@pytest.fixture()
def spec_kernels(monkeypatch):
    from allegro_b200 import _lib
    from allegro_b200.model.allegro_models import FusedAllegroEnergy

    for name in kernel_spec.ALL:
        monkeypatch.setattr(_lib, name, getattr(kernel_spec, name))
    monkeypatch.setattr(FusedAllegroEnergy, "core", lambda self: self._core_for(torch.device("cpu")))


def test_host_pipeline_with_pair_potential(spec_kernels, path="fused"):
    """Product host logic (CPU, kernels replaced by the executable spec) == oracle, energies / forces / stress, with the pair
    term added after the per-type scale/shift (the autograd forward is CUDA-only: covered by the GPU test below)."""
    from allegro_b200.model import AllegroModel

    d, avg = _frame()
    kw = _kw("float64", avg)
    kw["per_type_energy_scales"], kw["per_type_energy_shifts"] = [0.7, 1.3, 0.9], [0.1, -0.2, 0.3]
    oracle = AllegroOracle(**kw)
    model = AllegroModel(**kw)
    missing = model.load_state_dict(oracle.state_dict(), strict=True)
    assert "model.pair_potential.atomic_numbers" in model.state_dict()
    ref = oracle(d)
    if path == "autograd":
        model.use_autograd = True
        out = model(d)
    else:
        out = model.model._energy_and_forces(dict(d), True)
    for key in (D.PER_ATOM_ENERGY_KEY, D.TOTAL_ENERGY_KEY, D.FORCE_KEY) + ((D.STRESS_KEY,) if path == "fused" else ()):
        err = float((out[key] - ref[key]).abs().max() / ref[key].abs().max())
        assert err < 1e-10, (key, err)
    # the pair term matters in this frame (otherwise the test proves nothing)
    kw0 = dict(kw)
    kw0.pop("pair_potential")
    o0 = AllegroOracle(**kw0)
    o0.load_state_dict({k: v for k, v in oracle.state_dict().items() if "pair_potential" not in k})
    assert float((o0(d)[D.FORCE_KEY] - ref[D.FORCE_KEY]).abs().max()) > 1e-3 * float(ref[D.FORCE_KEY].abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["fused", "autograd"])
@pytest.mark.parametrize("dtype,tol", [("float64", 1e-9), ("float32", 1e-4)])
def test_gpu_model_with_pair_potential(dtype, tol, path):
    from allegro_b200.model import AllegroModel

    d, avg = _frame()
    kw = _kw("float64", avg)
    oracle = AllegroOracle(**kw)
    kwm = dict(kw)
    kwm["model_dtype"] = dtype
    model = AllegroModel(**kwm)
    model.load_state_dict(oracle.state_dict())
    model = model.to("cuda")
    model.use_autograd = path == "autograd"
    out = model({k: v.to("cuda") for k, v in d.items()})
    ref = oracle(d)
    for key in (D.PER_ATOM_ENERGY_KEY, D.FORCE_KEY) + ((D.STRESS_KEY,) if path == "fused" else ()):
        err = float((out[key].double().cpu() - ref[key]).abs().max() / ref[key].abs().max())
        assert err < tol, (key, err)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_gpu_zbl_kernel_matches_spec(dtype):
    from allegro_b200 import _lib

    g = torch.Generator().manual_seed(5)
    E, N, T = 4000, 300, 3
    vec = torch.randn(E, 3, generator=g, dtype=torch.float64) * 2.2
    vec[:50] *= 0.2  # short edges: the steep part of the screening function
    ctr, nbr = torch.randint(0, N, (E,), generator=g).int(), torch.randint(0, N, (E,), generator=g).int()
    types = torch.randint(0, T, (N,), generator=g).int()
    Z = torch.tensor([3.0, 15.0, 16.0], dtype=torch.float64)
    rmax = torch.tensor([[6.0, 5.0, 4.5], [5.0, 6.0, 5.5], [4.5, 5.5, 6.0]], dtype=torch.float64)
    g0 = torch.randn(E, 3, generator=g, dtype=torch.float64)
    want_g = g0.clone()
    want_e = kernel_spec.zbl(6.0, 7.1998225, vec, ctr, nbr, types, Z, rmax, want_g)
    dev = "cuda"
    got_g = g0.to(dev, dtype)
    got_e = _lib.zbl(6.0, 7.1998225, vec.to(dev, dtype), ctr.to(dev), nbr.to(dev), types.to(dev), Z.to(dev, dtype), rmax.to(dev, dtype), got_g)
    tol = 1e-12 if dtype == torch.float64 else 2e-5
    assert float((got_e.double().cpu() - want_e).abs().max() / want_e.abs().max()) < tol
    assert float((got_g.double().cpu() - want_g).abs().max() / want_g.abs().max()) < tol
