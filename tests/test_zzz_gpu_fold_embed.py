"""Opt-in host-side fusion (ALLEGRO_B200_FOLD_EMBED=1): the embed linears folded into the scalar-embed MLP's last layer.
Host logic is covered on the CPU (tests/test_host_pipeline.py); this runs the same thing through the CUDA kernels."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-9), ("float32", 1e-4)])
def test_folded_embed_matches_oracle(dtype, tol, monkeypatch):
    from test_gpu_model import _check, _pair

    monkeypatch.setenv("ALLEGRO_B200_FOLD_EMBED", "1")
    oracle, model, d = _pair("c2", 3, dtype)
    _check(oracle, model, d, tol, tol)
    assert model.model._upstream.fold
