"""A few launches of the kernels under ncu: L0 tensor-product fwd/bwd (c2 shapes) and two tensor-core linears."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from allegro_b200 import _lib, data as D
from allegro_b200.nn import Contracter

dev = "cuda"
N, deg, U, lmax = 10976, 42, 32, 2
E, Dd, nir = N * deg, 9, 3
ctr = torch.arange(N).repeat_interleave(deg)
csr = D.build_csr(torch.stack([ctr, (ctr + 1) % N]).to(dev), N)
dt = torch.float32
torch.manual_seed(0)
sh = "1x0e+1x1o+1x2e"
tp0 = Contracter(sh, sh, sh, mul=U)
ijk, _, _ = tp0.sparse_table()
tab, cgw = ijk.to(dev), tp0.cgw(dt, dev)
Y = torch.randn(E, Dd, device=dev); w0 = torch.randn(E, nir * U, device=dev); gam = torch.randn(N, Dd, U, device=dev)
Vout = torch.empty(E, 9, U, device=dev); go = torch.randn(E, 9, U, device=dev)
gw0 = torch.empty(E, nir * U, device=dev); gY = torch.zeros(E, Dd, device=dev); gg = torch.empty(N, Dd, U, device=dev)
for _ in range(2):
    _lib.tp_fwd(dt, lmax, N, E, U, 9, 9, tab, cgw, csr.row_ptr, csr.ctr, gam, None, Y, w0, Vout)
    _lib.tp_bwd(dt, lmax, N, E, U, 9, 9, tab, cgw, csr.row_ptr, csr.ctr, gam, None, Y, w0, go, None, gw0, gY, gg)
for awid, owid in (([64, 64, 32], [64]), ([64], [96, 64, 96])):
    a = [torch.randn(E, w, device=dev) for w in awid]
    W = torch.randn(sum(awid), sum(owid), device=dev) * 0.1
    o = [torch.zeros(E, w, device=dev) for w in owid]
    pk = _lib.linear_pack(W)
    _lib.linear(a, W, o, W_packed=pk)
torch.cuda.synchronize()
# environment adjoint (streaming) and the radial kernels
om = torch.randn(E, nir * U, device=dev); gom = torch.empty(E, nir * U, device=dev)
_lib.env_bwd(dt, lmax, U, csr.ctr, Y, om, gg, 0.15, gom, gY, row_ptr=csr.row_ptr)
_lib.env_sum(dt, lmax, N, U, csr.row_ptr, Y, om, 0.15)
vec = torch.randn(E, 3, device=dev) * 2.0
types = torch.zeros(N, dtype=torch.int32, device=dev)
rmax = torch.full((1, 1), 5.0, device=dev); bwv = torch.arange(1, 9, device=dev, dtype=torch.float32)
PQ = torch.randn(1, 8, 64, device=dev)
h = _lib.radial_pq_fwd(dt, 64, 6.0, vec, csr.ctr, csr.nbr, types, rmax, bwv, PQ)
gvec = torch.zeros(E, 3, device=dev)
_lib.radial_pq_bwd(dt, 64, 6.0, vec, csr.ctr, csr.nbr, types, rmax, bwv, PQ, torch.randn(E, 64, device=dev), h, gvec)
torch.cuda.synchronize()
