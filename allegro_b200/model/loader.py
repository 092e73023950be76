"""Checkpoint loader: reference / nequip ``state_dict`` -> ``allegro_b200.model.AllegroModel``.

Module names (hence key prefixes) are the reference's SequentialGraphNetwork keys
(/root/reference/allegro/model/allegro_models.py:222-228,262-268,297), so the Allegro-owned
keys (``allegro.tps.{l}.weights`` / ``.w3j``, ``per_type_energy_scale_shift.*`` ...) load by name.
Two things differ between producers and are handled here instead of silently dropped:

* ``ScalarMLPFunction`` lives in nequip, not in /root/reference, and its parameter names changed
  between nequip releases.  Inside every MLP prefix the weight matrices are therefore matched
  STRUCTURALLY: all >=2-D tensors under the prefix, in natural (numeric-aware) key order, are
  assigned to ``weights.{k}`` by position; each must have the layer's [h_in, h_out] shape
  (a torch.nn.Linear-style [h_out, h_in] ``*.weight`` is transposed).
* the dense ``w3j`` buffers are STATE (_contract.py:168): they are copied as data and the kernels
  build their sparse tables from the loaded values (``Contracter.sparse_table``), so a checkpoint
  produced with real e3nn (whatever its block signs) contracts with its own coupling tensor.
  A loaded ``w3j`` whose sparsity pattern differs from the selection rules raises.
"""
from __future__ import annotations

import re
from typing import Dict, List, Tuple

import torch

from ..nn._contract import Contracter
from ..nn._mlp import ScalarMLPFunction


def _natural(k: str):
    return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", k)]


def _strip_prefix(sd: Dict[str, torch.Tensor], model: torch.nn.Module) -> Dict[str, torch.Tensor]:
    """Drop wrapper prefixes (``model.``, ``sole_model.model.`` ...) so that keys start at the energy
    model's own sub-module names."""
    own_roots = {k.split(".")[0] for k in model.state_dict().keys()}
    out = {}
    for k, v in sd.items():
        parts = k.split(".")
        i = 0
        while i < len(parts) and parts[i] not in own_roots:
            i += 1
        out[".".join(parts[i:]) if i < len(parts) else k] = v
    return out


def map_reference_state_dict(model: torch.nn.Module, sd: Dict[str, torch.Tensor]) -> Tuple[Dict[str, torch.Tensor], List[str]]:
    """-> (state_dict with this package's key names, list of source keys that were not used)."""
    inner = getattr(model, "model", model)
    sd = _strip_prefix(dict(sd), inner)
    own = inner.state_dict()
    out: Dict[str, torch.Tensor] = {}
    used = set()
    # 1. MLPs: structural match inside each prefix
    for name, mod in inner.named_modules():
        if not isinstance(mod, ScalarMLPFunction):
            continue
        prefix = name + "."
        cands = sorted((k for k, v in sd.items() if k.startswith(prefix) and torch.is_tensor(v) and v.dim() >= 2), key=_natural)
        if len(cands) != len(mod.weights):
            raise KeyError(f"{name}: checkpoint has {len(cands)} weight matrices under this prefix, the model has {len(mod.weights)} layers")
        for k_idx, (src, w) in enumerate(zip(cands, mod.weights)):
            t = sd[src]
            t = t.reshape(t.shape[-2], t.shape[-1]) if t.dim() > 2 and t.numel() == t.shape[-2] * t.shape[-1] else t
            if tuple(t.shape) != tuple(w.shape):
                if tuple(t.T.shape) == tuple(w.shape) and (src.endswith(".weight") or t.shape[0] != t.shape[1]):
                    t = t.T
                else:
                    raise ValueError(f"{src}: shape {tuple(t.shape)} does not fit layer {k_idx} of {name} {tuple(w.shape)}")
            elif src.endswith(".weight") and t.shape[0] == t.shape[1]:
                t = t.T  # torch.nn.Linear convention [out, in]
            out[f"{prefix}weights.{k_idx}"] = t
            used.add(src)
    # 2. everything else by name
    for k in own:
        if k in out:
            continue
        if k in sd:
            out[k] = sd[k]
            used.add(k)
    return out, sorted(k for k in sd if k not in used)


def load_reference_state_dict(model: torch.nn.Module, sd: Dict[str, torch.Tensor], strict: bool = True):
    """Load a reference-produced ``state_dict`` and verify every loaded ``w3j`` against the selection
    rules (same non-zero pattern as the generated table; values -- signs, normalisation -- are the
    checkpoint's).  Returns the list of unused source keys."""
    inner = getattr(model, "model", model)
    mapped, unused = map_reference_state_dict(model, sd)
    res = inner.load_state_dict(mapped, strict=False)
    if strict and (res.missing_keys or res.unexpected_keys):
        raise KeyError(f"missing {res.missing_keys}, unexpected {res.unexpected_keys}")
    for name, mod in inner.named_modules():
        if isinstance(mod, Contracter):
            got = {(i, j, k, p) for i, j, k, p, _ in mod.w3j_entries()}
            want = {(i, j, k, p) for i, j, k, p, _ in mod.table.entries}
            if got != want:
                raise ValueError(f"{name}.w3j: loaded coupling tensor has {len(got ^ want)} non-zeros outside / missing from the "
                                 "O(3) selection rules of this layer -- wrong irreps or path order")
    return unused
