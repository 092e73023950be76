"""CPU tests pinning the oracle's (and the product's) O(3) arithmetic.

Restates what the reference pins through e3nn at run time
(/root/reference/tests/nn/test_contract_basic.py:120-211, tests/nn/test_weighter.py:12-54)
as known-answer + self-consistency checks, because e3nn is not installed here.
"""
import json
import math
import os

import numpy as np
import pytest
import torch

from allegro_b200 import o3
from oracle import o3_ref

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "o3_known_answers.json")))


def test_sh_known_answers():
    r = torch.tensor([[0.3, -0.5, 0.8]], dtype=torch.float64)
    Y = o3_ref.spherical_harmonics(2, r)[0]
    g = GOLD["sh_component_at_0.3_-0.5_0.8"]
    assert Y[0].item() == pytest.approx(1.0, abs=1e-12)
    np.testing.assert_allclose(Y[1:4].numpy(), g["l1"], atol=1e-9)
    np.testing.assert_allclose(Y[4:9].numpy(), g["l2"], atol=1e-9)


@pytest.mark.parametrize("lmax", [1, 2, 3, 5])
def test_sh_component_normalisation_and_recursion(lmax):
    pts = torch.randn(64, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    Y = o3_ref.spherical_harmonics(lmax, pts)
    for l in range(lmax + 1):
        n2 = (Y[:, l * l : (l + 1) ** 2] ** 2).sum(-1)
        np.testing.assert_allclose(n2.numpy(), 2 * l + 1, rtol=1e-12)
    if lmax <= 3:
        Yr = o3_ref.spherical_harmonics(lmax, pts, method="recursive")
        assert (Y - Yr).abs().max() < 1e-12
    # parity (-1)^l
    Ym = o3_ref.spherical_harmonics(lmax, -pts)
    for l in range(lmax + 1):
        sl = slice(l * l, (l + 1) ** 2)
        assert (Ym[:, sl] - (-1) ** l * Y[:, sl]).abs().max() < 1e-12


def test_w3j_known_answers():
    assert o3_ref.wigner_3j(1, 1, 1)[0, 1, 2] == pytest.approx(GOLD["w3j_111_012"], abs=1e-8)
    w = o3_ref.wigner_3j(1, 1, 2)
    nz = {tuple(int(v) for v in k.split(",")): val for k, val in GOLD["w3j_112_nonzeros"].items()}
    for idx in np.argwhere(w != 0):
        assert tuple(idx) in nz
    for idx, val in nz.items():
        assert w[idx] == pytest.approx(val, abs=1e-9)
    for key, n in GOLD["w3j_nnz"].items():
        l1, l2, l3 = (int(v) for v in key.split(","))
        assert int((o3_ref.wigner_3j(l1, l2, l3) != 0).sum()) == n
    for l in range(5):
        np.testing.assert_allclose(o3_ref.wigner_3j(l, l, 0)[:, :, 0], np.eye(2 * l + 1) / math.sqrt(2 * l + 1), atol=1e-14)


def test_sh_and_w3j_share_one_basis():
    pts = torch.randn(16, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    Y = o3_ref.spherical_harmonics(3, pts)
    sl = lambda l: slice(l * l, (l + 1) ** 2)  # noqa: E731
    for key, rho in GOLD["sh_product_rho"].items():
        a, b, c = (int(v) for v in key.split(","))
        w = torch.from_numpy(np.array(o3_ref.wigner_3j(a, b, c)))
        o = torch.einsum("ijk,zi,zj->zk", w, Y[:, sl(a)], Y[:, sl(b)])
        assert (o - rho * Y[:, sl(c)]).abs().max() < 2e-6


def test_product_w3j_equals_oracle_w3j():
    for l1 in range(5):
        for l2 in range(5):
            for l3 in range(abs(l1 - l2), min(l1 + l2, 5) + 1):
                a = np.array(o3.wigner_3j(l1, l2, l3))
                b = o3_ref.wigner_3j(l1, l2, l3)
                assert np.abs(a - b).max() < 1e-13


def test_w3j_is_an_intertwiner():
    """D1 x D2 x D3 leaves the 3j invariant (what e3nn's assert_equivariant checks indirectly)."""
    R = o3_ref.random_rotation(7)
    for l1, l2, l3 in [(1, 1, 2), (1, 2, 3), (2, 2, 2), (2, 3, 1), (3, 3, 3)]:
        D1, D2, D3 = (o3_ref.wigner_D_from_rotation(l, R) for l in (l1, l2, l3))
        w = torch.from_numpy(np.array(o3_ref.wigner_3j(l1, l2, l3)))
        w2 = torch.einsum("ai,bj,ck,ijk->abc", D1, D2, D3, w)
        assert (w - w2).abs().max() < 1e-10


def test_layer_tables_match_survey():
    for case in GOLD["layer_tables"]:
        lmax, L = case["l_max"], case["L"]
        sh = o3.Irreps.spherical_harmonics(lmax)
        allowed = o3.Irreps([(1, (l, p)) for l in range(lmax + 1) for p in (1, -1)])
        ins, outs = o3.allegro_layer_irreps(sh, allowed, L)
        ins_r, outs_r = __import__("oracle.nn_ref", fromlist=["x"]).allegro_layer_irreps(
            o3_ref.Irreps.spherical_harmonics(lmax),
            o3_ref.Irreps([(1, (l, p)) for l in range(lmax + 1) for p in (1, -1)]),
            L,
        )
        for a, b, ar, br, exp in zip(ins, outs, ins_r, outs_r, case["layers"]):
            t = o3.build_coupling_table(a, sh, b)
            assert (t.dim1, t.dim_out, t.num_paths, t.nnz, t.is_ij_diagonal) == (
                exp["d_in"], exp["d_out"], exp["P"], exp["nnz"], exp["diag"])
            assert repr(a) == repr(ar) and repr(b) == repr(br)


def test_irreps_parsing():
    a = o3.Irreps("0e + 0o + 1e + 1o")
    assert a.dim == 8 and a.num_irreps == 4 and a.lmax == 1
    assert repr(o3.Irreps("1x0e+1x1o")) == "1x0e+1x1o"
    assert o3.Irrep("2e") in o3.Irreps("2o + 1e + 2e")
    assert [repr(i) for i in o3.Irrep("1o") * o3.Irrep("2e")] == ["1o", "2o", "3o"]
    assert o3.Irreps.spherical_harmonics(2).comp_to_irrep() == [0, 1, 1, 1, 2, 2, 2, 2, 2]


def test_spherical_harmonics_against_scipy():
    """Independent implementation check: the oracle's (and the CUDA generator's) real spherical harmonics equal scipy's
    complex Y_l^m turned into the standard real basis, evaluated with e3nn's axis convention (y polar: physics
    (x, y, z) = e3nn (z, x, y)), component normalisation, m = -l..l."""
    import numpy as np
    import torch
    from scipy.special import sph_harm_y

    from oracle import o3_ref

    lmax = 4
    g = torch.Generator().manual_seed(3)
    v = torch.randn(400, 3, generator=g, dtype=torch.float64)
    v = v / v.norm(dim=-1, keepdim=True)
    mine = o3_ref.spherical_harmonics(lmax, v, method="recursive").numpy()
    x, y, z = v[:, 0].numpy(), v[:, 1].numpy(), v[:, 2].numpy()
    xp, yp, zp = z, x, y
    theta, phi = np.arccos(np.clip(zp, -1, 1)), np.arctan2(yp, xp)
    col = 0
    for l in range(lmax + 1):
        for m in range(-l, l + 1):
            c = sph_harm_y(l, abs(m), theta, phi)
            if m > 0:
                ref = np.sqrt(2) * (-1) ** m * c.real
            elif m < 0:
                ref = np.sqrt(2) * (-1) ** m * c.imag
            else:
                ref = c.real
            ref = ref * np.sqrt(4 * np.pi)
            a = mine[:, col]
            big = np.abs(ref) > 1e-3
            ratio = a[big] / ref[big]
            assert np.allclose(np.abs(ratio), 1.0, atol=1e-10), (l, m)
            assert np.all(ratio > 0), (l, m)  # no extra sign: exactly the standard real harmonics in e3nn's axes
            assert np.allclose(a, ref, atol=1e-10), (l, m)
            col += 1
    # explicit l <= 3 polynomials (the ones tools/gen_sh.py turns into CUDA) agree with the recursive construction
    assert np.allclose(o3_ref.spherical_harmonics(3, v, method="explicit").numpy(), mine[:, :16], atol=1e-12)


def test_wigner_3j_against_sympy_real_gaunt():
    """Independent check of the real-basis Wigner 3j: for l1+l2+l3 even it must be proportional to the integral of three
    real spherical harmonics (sympy's real Gaunt coefficients, an unrelated implementation), one constant per triple."""
    import numpy as np
    from sympy.physics.wigner import real_gaunt

    from oracle import o3_ref

    for l1, l2, l3 in [(0, 0, 0), (1, 1, 0), (1, 1, 2), (2, 1, 1), (2, 2, 2), (2, 2, 0), (1, 2, 3), (3, 3, 2), (2, 2, 4)]:
        w = np.array(o3_ref.wigner_3j(l1, l2, l3))
        gnt = np.zeros_like(w)
        for a in range(2 * l1 + 1):
            for b in range(2 * l2 + 1):
                for c in range(2 * l3 + 1):
                    gnt[a, b, c] = float(real_gaunt(l1, l2, l3, a - l1, b - l2, c - l3))
        assert (np.abs(gnt) > 1e-12).sum() == (w != 0).sum(), (l1, l2, l3)          # same sparsity pattern
        scale = (w * gnt).sum() / (gnt * gnt).sum()
        assert abs(scale) > 1e-6 and np.allclose(w, scale * gnt, atol=1e-12), (l1, l2, l3)
