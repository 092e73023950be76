"""Real-width parity ("parity only on toy sizes"): every BASELINE config at the feature widths it names --
c2 at its FULL size, c3 (S=128, U=64) and c5 (l_max=3, 3 layers, S=128, U=64, fp64) on reduced cells against the whole
oracle, and c3 at full size through the strict-locality sub-sample (oracle/subsample.py) plus size-independent
properties (zero net force, bitwise reproducibility)."""
import pytest
import torch

from allegro_b200 import data as D
from allegro_b200 import systems
from oracle.subsample import ball, local_reference
from test_gpu_model import _check, _pair, _to_dev

pytestmark = pytest.mark.gpu


def _sub_check(oracle, model, d, n_sample, tol, seed=0):
    dd = _to_dev(d)
    out = model(dd)
    atoms = ball(d[D.POSITIONS_KEY], n_sample, seed)
    centres, e_ref, f_ref = local_reference(oracle, d, atoms)
    e = out[D.PER_ATOM_ENERGY_KEY].double().cpu()[centres]
    f = out[D.FORCE_KEY].double().cpu()[atoms]
    err_e = float((e - e_ref).abs().max() / e_ref.abs().max())
    err_f = float((f - f_ref).abs().max() / f_ref.abs().max())
    assert err_e < tol and err_f < tol, (err_e, err_f)
    # size-independent properties on the whole frame
    F = out[D.FORCE_KEY].double()
    assert float(F.sum(0).abs().max()) < 1e-3 * tol * float(F.abs().sum(0).max()) + 1e-9 * float(F.abs().max())
    out2 = model(dd)
    core = model.model.core()
    if core.dtype == torch.float32 and core.U <= 32 and core.lmax <= 2:
        # every reduction on this path is a fixed-order segmented sum: bitwise reproducible
        assert torch.equal(out2[D.FORCE_KEY], out[D.FORCE_KEY])
        assert torch.equal(out2[D.PER_ATOM_ENERGY_KEY], out[D.PER_ATOM_ENERGY_KEY])
    else:
        # the shape-generic fp64 / multi-chunk kernels still accumulate ggamma / gY with atomics (order varies at the ulp level)
        eps = 1e-12 if core.dtype == torch.float64 else 1e-5
        assert float((out2[D.FORCE_KEY] - out[D.FORCE_KEY]).abs().max()) <= eps * float(out[D.FORCE_KEY].abs().max())
    return err_e, err_f


@pytest.mark.parametrize("dtype,tol", [("float32", 1e-4), ("float64", 1e-9)])
def test_c2_full_size(dtype, tol):
    """configs[1]: 10 976 atoms / 461k edges, l_max=2, 2 layers, S=64, U=32 -- the benchmark frame itself."""
    oracle, model, d = _pair("c2", None, dtype)
    assert d[D.POSITIONS_KEY].shape[0] == 10976
    ee, ef = _sub_check(oracle, model, d, 24, tol)
    print(f"c2 full {dtype}: E {ee:.2e} F {ef:.2e}")


@pytest.mark.parametrize("dtype,tol", [("float32", 1e-4), ("float64", 1e-9)])
def test_c3_real_width_reduced_cell(dtype, tol):
    """configs[2] architecture at its named widths (S=128, U=64, 3 species, r_max=6) on a 6^3 cell, whole oracle."""
    oracle, model, d = _pair("c3", 6, dtype)
    kw = systems.model_kwargs("c3", 1.0)
    assert kw["num_scalar_features"] == 128 and kw["num_tensor_features"] == 64
    _check(oracle, model, d, tol, tol)


def test_c5_real_width_fp64():
    """configs[4]: l_max=3, 3 layers, S=128, U=64, 5 species, fp64, on 2^3 FCC cells (32 atoms), whole oracle."""
    oracle, model, d = _pair("c5", 2, "float64")
    kw = systems.model_kwargs("c5", 1.0)
    assert kw["l_max"] == 3 and kw["num_layers"] == 3 and kw["num_tensor_features"] == 64
    _check(oracle, model, d, 1e-9, 1e-9)


def test_c5_full_size_fp64_subsample():
    """configs[4] at its full size (10 976 atoms, 5 species, fp64, full widths) through the locality sub-sample."""
    oracle, model, d = _pair("c5", None, "float64")
    ee, ef = _sub_check(oracle, model, d, 6, 1e-9)
    print(f"c5 full fp64: E {ee:.2e} F {ef:.2e}")


def test_c3_full_size_subsample():
    """configs[2] at full size: ~97k atoms, ~4.4M edges, S=128, U=64, fp32."""
    oracle, model, d = _pair("c3", None, "float32")
    assert d[D.POSITIONS_KEY].shape[0] > 90000
    ee, ef = _sub_check(oracle, model, d, 8, 1e-4)
    print(f"c3 full fp32: E {ee:.2e} F {ef:.2e}")
