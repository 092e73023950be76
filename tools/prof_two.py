"""A few launches under ncu: the last-layer (9 -> 1) streaming backward at the c2 shapes and the baked fp64 kernels at the c5 shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from allegro_b200 import _lib, data as D
from allegro_b200.nn import Contracter

dev = "cuda"
N, deg = 10976, 42
E = N * deg
ctr = torch.arange(N).repeat_interleave(deg)
csr = D.build_csr(torch.stack([ctr, (ctr + 1) % N]).to(dev), N)
torch.manual_seed(0)
# c2 last layer, fp32
sh2 = "1x0e+1x1o+1x2e"
tp = Contracter(sh2, sh2, "1x0e", mul=32)
ijk, _, _ = tp.sparse_table()
tab, cgw = ijk.to(dev), tp.cgw(torch.float32, dev)
Vin = torch.randn(E, 9, 32, device=dev); gam = torch.randn(N, 9, 32, device=dev); go = torch.randn(E, 1, 32, device=dev)
gVin = torch.empty(E, 9, 32, device=dev); gg = torch.empty(N, 9, 32, device=dev)
for _ in range(2):
    _lib.tp_bwd(torch.float32, 2, N, E, 32, 9, 1, tab, cgw, csr.row_ptr, csr.ctr, gam, Vin, None, None, go, gVin, None, None, gg)
torch.cuda.synchronize()
del Vin, gVin
# c5 layers 0 (implicit) and 1 (explicit), fp64, U = 64
U, Dd, nir = 64, 16, 4
sh3 = "1x0e+1x1o+1x2e+1x3o"
full = "1x0e+1x1e+1x1o+1x2e+1x2o+1x3e+1x3o"
prev = torch.get_default_dtype(); torch.set_default_dtype(torch.float64)
tp0, tp1 = Contracter(sh3, sh3, full, mul=U), Contracter(full, sh3, sh3, mul=U)
torch.set_default_dtype(prev)
dt = torch.float64
Y = torch.randn(E, Dd, device=dev, dtype=dt); w0 = torch.randn(E, nir * U, device=dev, dtype=dt); gam = torch.randn(N, Dd, U, device=dev, dtype=dt)
for tpx, implicit in ((tp0, True), (tp1, False)):
    ijk, _, _ = tpx.sparse_table()
    tab, cgw = ijk.to(dev), tpx.cgw(dt, dev)
    d_in, d_out = tpx.base_dim1, tpx.base_dim_out
    Vin = None if implicit else torch.randn(E, d_in, U, device=dev, dtype=dt)
    Vout = torch.empty(E, d_out, U, device=dev, dtype=dt); go = torch.randn(E, d_out, U, device=dev, dtype=dt)
    gVin = None if implicit else torch.empty(E, d_in, U, device=dev, dtype=dt)
    gw0 = torch.empty(E, nir * U, device=dev, dtype=dt) if implicit else None
    gY = torch.zeros(E, Dd, device=dev, dtype=dt) if implicit else None
    gg = torch.empty(N, Dd, U, device=dev, dtype=dt)
    _lib.tp_fwd(dt, 3, N, E, U, d_in, d_out, tab, cgw, csr.row_ptr, csr.ctr, gam, Vin, Y if implicit else None, w0 if implicit else None, Vout)
    _lib.tp_bwd(dt, 3, N, E, U, d_in, d_out, tab, cgw, csr.row_ptr, csr.ctr, gam, Vin, Y if implicit else None, w0 if implicit else None, go, gVin, gw0, gY, gg)
    torch.cuda.synchronize()
    del Vin, Vout, go, gVin, gg
