"""CUDA kernels against the executable specification of their C-ABI contract (tests/kernel_spec.py), for the entry points
that the older kernel suite only reaches through whole-model tests: ab2_edge_vec, ab2_radial_fwd / ab2_radial_bwd."""
import pytest
import torch

import kernel_spec
from allegro_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-11), (torch.float32, 1e-4)])
@pytest.mark.parametrize("S_rc,nb,T", [(16, 8, 1), (64, 8, 3), (32, 5, 5)])
def test_radial_fwd_bwd_and_edge_vec(dtype, tol, S_rc, nb, T):
    acc = _lib.ACC_DTYPE[dtype]
    g = torch.Generator().manual_seed(S_rc + nb + T)
    n_atoms, E = 40, 700
    pos = torch.randn(n_atoms, 3, generator=g, dtype=torch.float64) * 2.5
    ctr = torch.sort(torch.randint(0, n_atoms, (E,), generator=g)).values.to(torch.int32)
    nbr = torch.randint(0, n_atoms, (E,), generator=g).to(torch.int32)
    nbr = torch.where(nbr == ctr, (nbr + 1) % n_atoms, nbr).to(torch.int32)      # no zero-length edges
    shift = torch.randn(E, 3, generator=g, dtype=torch.float64) * 0.3
    types = torch.randint(0, T, (n_atoms,), generator=g).to(torch.int32)
    rmax = (3.0 + 2.0 * torch.rand(T, T, generator=g, dtype=torch.float64)).to(acc)   # some edges beyond, some inside
    bw = torch.linspace(1.0, nb, nb, dtype=torch.float64).to(acc)
    Wb = (torch.randn(nb, S_rc, generator=g, dtype=torch.float64) / nb**0.5).to(acc)
    cemb = torch.randn(T, S_rc // 2, generator=g, dtype=torch.float64).to(acc)
    nemb = torch.randn(T, S_rc // 2, generator=g, dtype=torch.float64).to(acc)
    p_cut = 6.0

    def dev(t):
        return t.to(DEV)

    vec_ref = kernel_spec.edge_vec(pos, ctr, nbr, shift, acc)
    vec = _lib.edge_vec(dev(pos), dev(ctr), dev(nbr), dev(shift), acc)
    assert (vec.cpu() - vec_ref).abs().max() < (1e-12 if acc == torch.float64 else 1e-5)
    assert (_lib.edge_vec(dev(pos), dev(ctr), dev(nbr), None, acc).cpu() - kernel_spec.edge_vec(pos, ctr, nbr, None, acc)).abs().max() < 1e-5

    # reference in fp64 arithmetic on the (possibly fp32-rounded) inputs the kernel sees
    args_ref = (torch.float64, S_rc, p_cut, vec_ref.double(), ctr, nbr, types, rmax.double(), bw.double(), Wb.double(), cemb.double(), nemb.double())
    args_dev = (dtype, S_rc, p_cut, dev(vec_ref), dev(ctr), dev(nbr), dev(types), dev(rmax), dev(bw), dev(Wb), dev(cemb), dev(nemb))
    e0_ref = kernel_spec.radial_fwd(*args_ref)
    e0 = _lib.radial_fwd(*args_dev)
    scale = float(e0_ref.abs().max())
    assert (e0.cpu().double() - e0_ref.double()).abs().max() < tol * scale
    x = vec_ref.norm(dim=-1) / rmax[types.long()[ctr.long()], types.long()[nbr.long()]]
    assert bool((x >= 1).any()) and bool((x < 1).any())
    assert float(e0.cpu()[x >= 1].abs().max()) == 0.0               # beyond the (per-type) cutoff: exactly zero

    g_e0 = torch.randn(E, S_rc, generator=g, dtype=torch.float64).to(dtype)
    gvec0 = torch.randn(E, 3, generator=g, dtype=torch.float64).to(acc)
    gv_ref = gvec0.double().clone()
    kernel_spec.radial_bwd(*args_ref, g_e0.double(), gv_ref)
    gv = dev(gvec0.clone())
    _lib.radial_bwd(*args_dev, dev(g_e0), gv)                           # accumulates into gvec
    assert (gv.cpu().double() - gv_ref.double()).abs().max() < tol * float(gv_ref.abs().max()) * 10


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-11), (torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("S,T,with_aux", [(64, 1, True), (128, 3, False), (16, 5, True), (100, 2, False)])
def test_radial_pq_fwd_bwd(dtype, tol, S, T, with_aux):
    """ab2_radial_pq_fwd / bwd (per-type-pair matrices, optional silu' factor in the adjoint) against the specification."""
    acc = _lib.ACC_DTYPE[dtype]
    nb = 8
    g = torch.Generator().manual_seed(S + T)
    n_atoms, E = 50, 1000      # > 3 blocks of 256 edges, last one partial
    ctr = torch.sort(torch.randint(0, n_atoms, (E,), generator=g)).values.to(torch.int32)
    nbr = torch.randint(0, n_atoms, (E,), generator=g).to(torch.int32)
    types = torch.randint(0, T, (n_atoms,), generator=g).to(torch.int32)
    vec = (torch.randn(E, 3, generator=g, dtype=torch.float64) * 2.0).to(acc)
    rmax = (3.0 + 2.0 * torch.rand(T, T, generator=g, dtype=torch.float64)).to(acc)
    bw = torch.linspace(1.0, nb, nb, dtype=torch.float64).to(acc)
    PQ = (torch.randn(T * T, nb, S, generator=g, dtype=torch.float64) / nb**0.5).to(acc)
    p_cut = 6.0

    def dev(t):
        return None if t is None else t.to(DEV)

    ref = kernel_spec.radial_pq_fwd(torch.float64, S, p_cut, vec.double(), ctr, nbr, types, rmax.double(), bw.double(), PQ.double())
    out = _lib.radial_pq_fwd(dtype, S, p_cut, dev(vec), dev(ctr), dev(nbr), dev(types), dev(rmax), dev(bw), dev(PQ))
    assert (out.cpu().double() - ref).abs().max() < tol * float(ref.abs().max())
    g_out = torch.randn(E, S, generator=g, dtype=torch.float64).to(dtype)
    aux = torch.randn(E, S, generator=g, dtype=torch.float64).to(dtype) if with_aux else None
    gvec0 = torch.randn(E, 3, generator=g, dtype=torch.float64).to(acc)
    gv_ref = gvec0.double().clone()
    kernel_spec.radial_pq_bwd(torch.float64, S, p_cut, vec.double(), ctr, nbr, types, rmax.double(), bw.double(), PQ.double(), g_out.double(),
                              None if aux is None else aux.double(), gv_ref)
    gv = dev(gvec0.clone())
    _lib.radial_pq_bwd(dtype, S, p_cut, dev(vec), dev(ctr), dev(nbr), dev(types), dev(rmax), dev(bw), dev(PQ), dev(g_out), dev(aux), gv)
    assert (gv.cpu().double() - gv_ref).abs().max() < (tol if dtype != torch.bfloat16 else 1e-4) * float(gv_ref.abs().max()) * 10
