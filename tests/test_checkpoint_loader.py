"""Checkpoint fidelity (VERDICT r1 item 8 / ADVICE): the loaded dense ``w3j`` buffer is what the kernels
contract with, nequip-style MLP key names map structurally, and the implicit per-neighbour-list caches cannot alias
another frame.  Host logic only (kernels replaced by tests/kernel_spec.py)."""
import pytest
import torch

from golden_util import load_models, unpack_state_dict
from test_host_pipeline import _rel, spec_kernels  # noqa: F401  (fixture)

MODELS = {r["name"]: r for r in load_models()}


def _pair(name):
    from allegro_b200.model import AllegroModel
    from oracle.model_ref import AllegroOracle

    rec = MODELS[name]
    sd = unpack_state_dict(rec["state_dict"])
    model = AllegroModel(**rec["kwargs"])
    oracle = AllegroOracle(**rec["kwargs"])
    return rec, sd, model, oracle


def test_loaded_w3j_is_used_by_the_kernels(spec_kernels):
    """Flip the sign of one (l1,l2,l3) block and rescale another in the checkpoint's w3j: the product must follow the
    checkpoint (as the reference's dense einsum over the buffer does, _contract.py:218-219), not its own table."""
    rec, sd, model, oracle = _pair("c5_lmax3_L3_5species")
    key = "model.allegro.tps.0.w3j"
    w = sd[key].clone()
    assert w.dim() == 4
    w[1] = -w[1]            # an e3nn build with the opposite sign convention on this path
    w[2] = 1.25 * w[2]      # and a different normalisation on that one
    sd[key] = w
    oracle.load_state_dict(sd, strict=True)
    model.load_state_dict(sd, strict=True)
    ref = oracle(dict(rec["data"]))
    out = model.model._energy_and_forces(dict(rec["data"]), False)
    assert _rel(out["forces"], ref["forces"]) < 1e-10 and _rel(out["atomic_energy"], ref["atomic_energy"]) < 1e-10
    # and it is a different model from the unmodified checkpoint
    assert _rel(out["forces"], rec["forces"]) > 1e-3


def test_reloading_weights_rebuilds_the_core(spec_kernels):
    rec, sd, model, oracle = _pair("c2_lmax2_L2")
    model.load_state_dict(sd, strict=True)
    out1 = model.model._energy_and_forces(dict(rec["data"]), False)
    sd2 = {k: (v * 1.1 if k.endswith("tps.1.weights") else v) for k, v in sd.items()}
    model.load_state_dict(sd2, strict=True)
    out2 = model.model._energy_and_forces(dict(rec["data"]), False)
    oracle.load_state_dict(sd2, strict=True)
    assert _rel(out2["forces"], oracle(dict(rec["data"]))["forces"]) < 1e-10
    assert _rel(out2["forces"], out1["forces"]) > 1e-6


@pytest.mark.parametrize("style", ["nequip_mlp_layers", "linear_weight_T", "wrapped_prefix"])
def test_mlp_key_map(style, spec_kernels):
    """nequip's ScalarMLPFunction parameter names are not the ones this package uses; the loader matches the weight
    matrices of every MLP prefix by position and shape."""
    from allegro_b200.model.loader import load_reference_state_dict

    rec, sd, model, oracle = _pair("c2_lmax2_L2")
    ren = {}
    for k, v in sd.items():
        if ".weights." in k and v.dim() == 2:
            pre, idx = k.rsplit(".weights.", 1)
            if style == "nequip_mlp_layers":
                ren[f"{pre}.mlp.layer_{idx}.weights"] = v
            elif style == "linear_weight_T":
                ren[f"{pre}.layers.{idx}.weight"] = v.T.contiguous()
            else:
                ren[f"sole_model.{pre}._weight_{idx}"] = v
        else:
            ren[("sole_model." + k) if style == "wrapped_prefix" else k] = v
    unused = load_reference_state_dict(model, ren)
    assert unused == []
    out = model.model._energy_and_forces(dict(rec["data"]), False)
    assert _rel(out["forces"], rec["forces"]) < 1e-10


def test_loader_rejects_a_w3j_with_the_wrong_pattern():
    from allegro_b200.model.loader import load_reference_state_dict

    rec, sd, model, _ = _pair("c2_lmax2_L2")
    w = sd["model.allegro.tps.0.w3j"].clone()
    w[0, 0, 1, 0] = 0.5  # (l1,l2,l3) = (0,1,0): forbidden by the selection rules
    sd["model.allegro.tps.0.w3j"] = w
    with pytest.raises(ValueError, match="selection rules"):
        load_reference_state_dict(model, sd)


def test_caches_do_not_alias_a_new_frame(spec_kernels):
    """Two frames whose tensors share shape (and, on a GPU, possibly the address): the second must not reuse the first
    frame's CSR / types / shifts."""
    from allegro_b200 import data as D

    rec, sd, model, oracle = _pair("c5_lmax3_L3_5species")
    model.load_state_dict(sd, strict=True)
    oracle.load_state_dict(sd, strict=True)
    d1 = dict(rec["data"])
    out1 = model.model._energy_and_forces(d1, False)
    # same atoms, types permuted, edges re-derived after a displacement -> same shapes, different content
    g = torch.Generator().manual_seed(3)
    pos = d1[D.POSITIONS_KEY] + 0.3 * torch.randn(d1[D.POSITIONS_KEY].shape, generator=g, dtype=d1[D.POSITIONS_KEY].dtype)
    ei, sh = D.neighbor_list(pos, rec["kwargs"]["r_max"], d1[D.CELL_KEY], (True, True, True))
    types = d1[D.ATOM_TYPE_KEY].flip(0).contiguous()
    d2 = {D.POSITIONS_KEY: pos, D.CELL_KEY: d1[D.CELL_KEY], D.ATOM_TYPE_KEY: types, D.EDGE_INDEX_KEY: ei, D.EDGE_CELL_SHIFT_KEY: sh}
    out2 = model.model._energy_and_forces(d2, False)
    ref2 = oracle(d2)
    assert _rel(out2["forces"], ref2["forces"]) < 1e-10 and _rel(out2["atomic_energy"], ref2["atomic_energy"]) < 1e-10
    # in-place edit of a cached source tensor is seen (version counter)
    d2[D.ATOM_TYPE_KEY].copy_(d1[D.ATOM_TYPE_KEY])
    out3 = model.model._energy_and_forces(d2, False)
    assert _rel(out3["atomic_energy"], oracle(d2)["atomic_energy"]) < 1e-10
    del out1


def test_batched_frames_are_rejected(spec_kernels):
    from allegro_b200 import data as D

    rec, sd, model, _ = _pair("c2_lmax2_L2")
    d = dict(rec["data"])
    n = d[D.POSITIONS_KEY].shape[0]
    d[D.BATCH_KEY] = torch.cat([torch.zeros(n // 2, dtype=torch.long), torch.ones(n - n // 2, dtype=torch.long)])
    with pytest.raises(NotImplementedError, match="batched"):
        model.model._energy_and_forces(d, False)


def test_neighbor_list_wraps_unwrapped_positions():
    """the all-pairs path must see the same neighbours when atoms sit in other periodic images."""
    from allegro_b200 import data as D

    g = torch.Generator().manual_seed(0)
    cell = torch.tensor([[6.0, 0, 0], [1.0, 5.5, 0], [0.3, 0.2, 6.2]], dtype=torch.float64)
    pos = torch.rand(20, 3, generator=g, dtype=torch.float64) @ cell
    ei, sh = D.neighbor_list(pos, 3.0, cell)
    pos2 = pos.clone()
    pos2[3] += 2 * cell[0] - cell[2]
    pos2[7] -= 3 * cell[1]
    pos2[11] += 11 * cell[2]
    ei2, sh2 = D.neighbor_list(pos2, 3.0, cell)
    assert torch.equal(ei, ei2)
    v1 = pos[ei[1]] + sh @ cell - pos[ei[0]]
    v2 = pos2[ei2[1]] + sh2 @ cell - pos2[ei2[0]]
    assert (v1 - v2).abs().max() < 1e-12 and float(v1.norm(dim=-1).max()) < 3.0
    # orthorhombic: brute force and cell list agree on raw positions far from the home cell
    box = torch.diag(torch.tensor([16.0, 15.5, 17.0], dtype=torch.float64))
    p = torch.rand(300, 3, generator=g, dtype=torch.float64) @ box
    p[5] += 3 * box[0]
    p[17] -= 2 * box[2]
    a, sa = D.neighbor_list(p, 4.0, box, method="brute")
    b, sb = D.neighbor_list(p, 4.0, box, method="cell")
    assert torch.equal(a, b) and torch.equal(sa, sb)
