"""Loading helpers for the reference-generated fixtures under tests/golden/ (see
tests/golden/make_reference_vectors.py for how they were produced)."""
import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def unpack_state_dict(sd):
    """Inverse of make_reference_vectors.pack_state_dict: densify the (shape, idx, val) `w3j` records."""
    out = {}
    for k, v in sd.items():
        if isinstance(v, dict):
            t = torch.zeros(v["w3j_shape"], dtype=v["dtype"])
            t[tuple(v["idx"].long().T)] = v["val"]
            out[k] = t
        else:
            out[k] = v
    return out


def load_models():
    return torch.load(os.path.join(GOLDEN, "ref_models.pt"), weights_only=False)


def load_ops():
    return torch.load(os.path.join(GOLDEN, "ref_ops.pt"), weights_only=False)


def model_case_ids():
    return [r["name"] for r in load_models()]
