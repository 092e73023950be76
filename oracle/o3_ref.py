"""Oracle O(3) arithmetic (test infrastructure; see oracle/__init__.py).

Restates the published e3nn algorithms the reference calls but does not vendor:

* ``wigner_3j``           <- e3nn.o3._wigner.wigner_3j, called at
                             allegro/nn/_strided/_contract.py:95
* ``spherical_harmonics`` <- e3nn.o3._spherical_harmonics.SphericalHarmonics,
                             built at allegro/nn/tensorembed.py:55-57, called :92
* ``Irrep`` / ``Irreps``  <- e3nn.o3._irreps, used at allegro/nn/_allegro.py:58,101-160
                             and allegro/nn/_strided/_contract.py:56-72

numpy float64 / complex128 throughout; torch only for the differentiable SH.
"""
from __future__ import annotations

import math
import re
from fractions import Fraction
from functools import lru_cache
from typing import Iterable, List, Sequence, Tuple, Union

import numpy as np
import torch


# --------------------------------------------------------------------------- #
# Irreps bookkeeping
# --------------------------------------------------------------------------- #
class Irrep(tuple):
    """(l, p) with p = +1 (even, 'e') or -1 (odd, 'o')."""

    def __new__(cls, l, p=None):
        if p is None:
            if isinstance(l, Irrep):
                return l
            if isinstance(l, str):
                m = re.fullmatch(r"\s*(\d+)([eo])\s*", l)
                assert m, f"bad irrep {l!r}"
                l, p = int(m.group(1)), (1 if m.group(2) == "e" else -1)
            else:
                l, p = l
        assert l >= 0 and p in (1, -1)
        return super().__new__(cls, (int(l), int(p)))

    @property
    def l(self):  # noqa: E743
        return self[0]

    @property
    def p(self):
        return self[1]

    @property
    def dim(self):
        return 2 * self[0] + 1

    def __mul__(self, other):
        other = Irrep(other)
        p = self.p * other.p
        return [Irrep(l, p) for l in range(abs(self.l - other.l), self.l + other.l + 1)]

    def __repr__(self):
        return f"{self.l}{'e' if self.p == 1 else 'o'}"


class Irreps(tuple):
    """Ordered list of (mul, Irrep)."""

    def __new__(cls, spec=None):
        if isinstance(spec, Irreps):
            return spec
        out = []
        if spec is None:
            spec = []
        if isinstance(spec, str):
            for term in spec.split("+"):
                term = term.strip()
                if not term:
                    continue
                if "x" in term:
                    mul, ir = term.split("x")
                    out.append((int(mul), Irrep(ir)))
                else:
                    out.append((1, Irrep(term)))
        else:
            for item in spec:
                if isinstance(item, (Irrep, str)):
                    out.append((1, Irrep(item)))
                else:
                    mul, ir = item
                    out.append((int(mul), Irrep(ir)))
        return super().__new__(cls, out)

    @staticmethod
    def spherical_harmonics(lmax: int, p: int = -1) -> "Irreps":
        return Irreps([(1, (l, p**l)) for l in range(lmax + 1)])

    @property
    def dim(self):
        return sum(mul * ir.dim for mul, ir in self)

    @property
    def num_irreps(self):
        return sum(mul for mul, _ in self)

    @property
    def lmax(self):
        return max(ir.l for _, ir in self)

    def slices(self):
        out, i = [], 0
        for mul, ir in self:
            out.append(slice(i, i + mul * ir.dim))
            i += mul * ir.dim
        return out

    def __contains__(self, ir):
        ir = Irrep(ir)
        return any(i == ir for _, i in self)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return Irreps(tuple.__getitem__(self, i))
        return tuple.__getitem__(self, i)

    def __repr__(self):
        return "+".join(f"{mul}x{ir}" for mul, ir in self)

    def randn(self, *size, generator=None, dtype=None):
        size = [self.dim if s == -1 else s for s in size]
        return torch.randn(*size, generator=generator, dtype=dtype)


# --------------------------------------------------------------------------- #
# Wigner 3j in e3nn's real basis
# --------------------------------------------------------------------------- #
def _f(n: int) -> int:
    return math.factorial(round(n))


def _su2_cg_coeff(j1, m1, j2, m2, j3, m3) -> float:
    """<j1 m1 j2 m2 | j3 m3>, Racah's closed form (integer j only is enough here)."""
    if m3 != m1 + m2:
        return 0.0
    vmin = int(max(-j1 + j2 + m3, -j1 + m1, 0))
    vmax = int(min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3))
    C = (
        (2.0 * j3 + 1.0)
        * Fraction(
            _f(j3 + j1 - j2) * _f(j3 - j1 + j2) * _f(j1 + j2 - j3) * _f(j3 + m3) * _f(j3 - m3),
            _f(j1 + j2 + j3 + 1) * _f(j1 - m1) * _f(j1 + m1) * _f(j2 - m2) * _f(j2 + m2),
        )
    ) ** 0.5
    S = 0
    for v in range(vmin, vmax + 1):
        S += (-1) ** int(v + j2 + m2) * Fraction(
            _f(j2 + j3 + m1 - v) * _f(j1 - m1 + v),
            _f(v) * _f(j3 - j1 + j2 - v) * _f(j3 + m3 - v) * _f(v + j1 - j2 - m3),
        )
    return float(C * S)


def _su2_cg(j1: int, j2: int, j3: int) -> np.ndarray:
    mat = np.zeros((2 * j1 + 1, 2 * j2 + 1, 2 * j3 + 1))
    if abs(j1 - j2) <= j3 <= j1 + j2:
        for m1 in range(-j1, j1 + 1):
            for m2 in range(-j2, j2 + 1):
                if abs(m1 + m2) <= j3:
                    mat[j1 + m1, j2 + m2, j3 + m1 + m2] = _su2_cg_coeff(j1, m1, j2, m2, j3, m1 + m2)
    return mat


def _real_to_complex(l: int) -> np.ndarray:
    """e3nn's change of basis real -> complex SH, incl. the (-i)^l phase that makes CG real."""
    q = np.zeros((2 * l + 1, 2 * l + 1), dtype=np.complex128)
    s = 1 / math.sqrt(2)
    for m in range(-l, 0):
        q[l + m, l + abs(m)] = s
        q[l + m, l - abs(m)] = -1j * s
    q[l, l] = 1
    for m in range(1, l + 1):
        q[l + m, l + abs(m)] = (-1) ** m * s
        q[l + m, l - abs(m)] = 1j * (-1) ** m * s
    return (-1j) ** l * q


@lru_cache(maxsize=None)
def wigner_3j(l1: int, l2: int, l3: int) -> np.ndarray:
    """Real-basis Wigner 3j, Frobenius norm 1, shape (2l1+1, 2l2+1, 2l3+1)."""
    assert abs(l2 - l3) <= l1 <= l2 + l3
    Q1, Q2, Q3 = _real_to_complex(l1), _real_to_complex(l2), _real_to_complex(l3)
    C = _su2_cg(l1, l2, l3).astype(np.complex128)
    C = np.einsum("ij,kl,mn,ikn->jlm", Q1, Q2, np.conj(Q3.T), C)
    assert np.abs(C.imag).max() < 1e-10
    C = C.real
    C = C / np.linalg.norm(C)
    C[np.abs(C) < 1e-14] = 0.0
    C.setflags(write=False)
    return C


# --------------------------------------------------------------------------- #
# Spherical harmonics, e3nn convention (y is the polar axis, m = -l..l)
# --------------------------------------------------------------------------- #
def _sh_explicit(lmax: int, x, y, z) -> List[torch.Tensor]:
    """Component-normalised polynomials for unit (x,y,z), l <= 3."""
    out = [torch.ones_like(x)]
    if lmax >= 1:
        s3 = math.sqrt(3.0)
        out += [s3 * x, s3 * y, s3 * z]
    if lmax >= 2:
        s15, s5 = math.sqrt(15.0), math.sqrt(5.0)
        x2, y2, z2 = x * x, y * y, z * z
        sh20 = s15 * x * z
        sh24 = 0.5 * s15 * (z2 - x2)
        out += [sh20, s15 * x * y, s5 * (y2 - 0.5 * (x2 + z2)), s15 * y * z, sh24]
    if lmax >= 3:
        q = x2 + z2
        c0 = math.sqrt(42.0) / 6.0
        c1 = math.sqrt(7.0)
        c2 = math.sqrt(168.0) / 8.0
        out += [
            c0 * (sh20 * z + sh24 * x),
            c1 * sh20 * y,
            c2 * (4.0 * y2 - q) * x,
            0.5 * c1 * y * (2.0 * y2 - 3.0 * q),
            c2 * z * (4.0 * y2 - q),
            c1 * sh24 * y,
            c0 * (sh24 * z - sh20 * x),
        ]
    return out


def _sh_recursive(lmax: int, x, y, z) -> List[torch.Tensor]:
    """Any lmax: Y_{l} = c_l * w3j(l-1,1,l)[i,j,k] Y_{l-1,i} Y_{1,j}, c_l > 0 fixed by
    sum_m Y_lm^2 = 2l+1 (how e3nn generates its polynomials)."""
    s3 = math.sqrt(3.0)
    y1 = torch.stack([s3 * x, s3 * y, s3 * z], dim=-1)
    Ys = [torch.ones_like(x).unsqueeze(-1), y1]
    north = torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64)
    ref = [torch.ones(1, dtype=torch.float64), s3 * north]
    for l in range(2, lmax + 1):
        w = torch.from_numpy(np.array(wigner_3j(l - 1, 1, l))).to(dtype=x.dtype, device=x.device)
        raw = torch.einsum("ijk,...i,...j->...k", w, Ys[l - 1], y1)
        rawref = torch.einsum("ijk,i,j->k", w.double().cpu(), ref[l - 1], ref[1])
        c = math.sqrt(2 * l + 1) / float(rawref.norm())
        Ys.append(c * raw)
        ref.append(c * rawref)
    flat = []
    for l in range(lmax + 1):
        flat += list(Ys[l].unbind(-1))
    return flat[: (lmax + 1) ** 2]


def spherical_harmonics(
    lmax: int,
    vec: torch.Tensor,
    normalize: bool = True,
    normalization: str = "component",
    method: str = "auto",
) -> torch.Tensor:
    """[..., 3] -> [..., (lmax+1)^2]; e3nn o3.SphericalHarmonics(0..lmax, normalize, normalization)."""
    if normalize:
        vec = vec / vec.norm(dim=-1, keepdim=True)
    x, y, z = vec.unbind(-1)
    if method == "auto":
        method = "explicit" if lmax <= 3 else "recursive"
    comps = _sh_explicit(lmax, x, y, z) if method == "explicit" else _sh_recursive(lmax, x, y, z)
    out = torch.stack(comps, dim=-1)
    if normalization == "component":
        return out
    scale = []
    for l in range(lmax + 1):
        f = 1.0 / math.sqrt(2 * l + 1) if normalization == "norm" else 1.0 / math.sqrt(4 * math.pi)
        scale += [f] * (2 * l + 1)
    assert normalization in ("norm", "integral")
    return out * torch.tensor(scale, dtype=out.dtype, device=out.device)


# --------------------------------------------------------------------------- #
# Wigner D in the same real basis (for equivariance tests)
# --------------------------------------------------------------------------- #
def wigner_D_from_rotation(l: int, R: torch.Tensor) -> torch.Tensor:
    """D^l(R) in the SH basis above, solved by least squares from Y_l(R r) = D Y_l(r)."""
    g = torch.Generator().manual_seed(1000 + l)
    pts = torch.randn(8 * (2 * l + 1), 3, generator=g, dtype=torch.float64)
    sl = slice(l * l, (l + 1) * (l + 1))
    A = spherical_harmonics(l, pts, method="recursive" if l > 3 else "explicit")[:, sl]
    B = spherical_harmonics(l, pts @ R.double().T, method="recursive" if l > 3 else "explicit")[:, sl]
    # B = A @ D^T
    Dt = torch.linalg.lstsq(A, B).solution
    return Dt.T


def random_rotation(seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    q, r = torch.linalg.qr(torch.randn(3, 3, generator=g, dtype=torch.float64))
    q = q * torch.sign(torch.diagonal(r))
    if torch.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q
