"""nequip.data.AtomicDataDict: field-name constants and the two helpers the reference calls
(num_nodes: allegro/nn/_allegro.py:239, allegro/nn/edgewise.py:56)."""
from typing import Dict

import torch

from oracle import nn_ref as _R

Type = Dict[str, torch.Tensor]

POSITIONS_KEY = _R.POSITIONS_KEY
EDGE_INDEX_KEY = _R.EDGE_INDEX_KEY
ATOM_TYPE_KEY = _R.ATOM_TYPE_KEY
CELL_KEY = _R.CELL_KEY
EDGE_CELL_SHIFT_KEY = _R.EDGE_CELL_SHIFT_KEY
EDGE_VECTORS_KEY = _R.EDGE_VECTORS_KEY
EDGE_LENGTH_KEY = _R.EDGE_LENGTH_KEY
NORM_LENGTH_KEY = _R.NORM_LENGTH_KEY
EDGE_TYPE_KEY = _R.EDGE_TYPE_KEY
EDGE_ATTRS_KEY = _R.EDGE_ATTRS_KEY
EDGE_EMBEDDING_KEY = _R.EDGE_EMBEDDING_KEY
EDGE_FEATURES_KEY = _R.EDGE_FEATURES_KEY
EDGE_ENERGY_KEY = _R.EDGE_ENERGY_KEY
EDGE_CUTOFF_KEY = "edge_cutoff"
PER_ATOM_ENERGY_KEY = _R.PER_ATOM_ENERGY_KEY
TOTAL_ENERGY_KEY = _R.TOTAL_ENERGY_KEY
FORCE_KEY = _R.FORCE_KEY


def num_nodes(data: Type) -> int:
    return data[POSITIONS_KEY].shape[0]
