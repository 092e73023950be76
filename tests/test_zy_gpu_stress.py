"""Stress / virial of the CUDA path (from the per-edge gradients, energy_and_forces(stress=True)) against the
oracle's strain derivative (nequip ForceStressOutput restated).  Sorted after the parity suites on purpose."""
import pytest
import torch

from allegro_b200 import data as D

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg,scale,dtype,tol", [("c2", 3, "float64", 1e-9), ("c5", 2, "float64", 1e-9), ("c2", 3, "float32", 1e-4)])
def test_stress_and_virial(cfg, scale, dtype, tol):
    from test_gpu_model import SMALL, _pair, _to_dev

    over = dict(SMALL) if cfg == "c5" else {}
    oracle, model, d = _pair(cfg, scale, dtype, **over)
    ref = oracle(d)
    out = model(_to_dev(d))
    assert out[D.STRESS_KEY].shape == (1, 3, 3) and out[D.VIRIAL_KEY].shape == (1, 3, 3)
    s_ref, s = ref[D.STRESS_KEY][0], out[D.STRESS_KEY][0].double().cpu()
    assert (s - s_ref).abs().max() / s_ref.abs().max() < tol
    v_ref, v = ref[D.VIRIAL_KEY][0], out[D.VIRIAL_KEY][0].double().cpu()
    assert (v - v_ref).abs().max() / v_ref.abs().max() < tol
    # forces and energies are untouched by the extra output
    assert (out[D.FORCE_KEY].double().cpu() - ref[D.FORCE_KEY]).abs().max() / ref[D.FORCE_KEY].abs().max() < max(tol, 1e-9)
    # opt-out: no stress keys, same forces
    model.compute_stress = False
    out2 = model(_to_dev(d))
    assert D.STRESS_KEY not in out2
    assert (out2[D.FORCE_KEY] - out[D.FORCE_KEY]).abs().max() <= 1e-6 * out[D.FORCE_KEY].abs().max()  # atomics: not bitwise
