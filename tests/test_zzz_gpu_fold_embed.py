"""Host-side fusion of the embed linears into the scalar-embed MLP's last layer (default on; ALLEGRO_B200_FOLD_EMBED=0 turns it
off).  Host logic is covered on the CPU (tests/test_host_pipeline.py); this runs both settings through the CUDA kernels."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fold", ["1", "0"])
@pytest.mark.parametrize("dtype,tol", [("float64", 1e-9), ("float32", 1e-4)])
def test_folded_embed_matches_oracle(dtype, tol, fold, monkeypatch):
    from test_gpu_model import _check, _pair

    monkeypatch.setenv("ALLEGRO_B200_FOLD_EMBED", fold)
    oracle, model, d = _pair("c2", 3, dtype)
    _check(oracle, model, d, tol, tol)
    assert model.model._upstream.fold == (fold == "1")
