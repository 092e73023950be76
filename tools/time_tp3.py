"""Stage knock-outs of the three-warp layer-0 backward (tp_stream3.cu) at the c2 shapes: which part bounds it?"""
import sys, os
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
go = torch.randn(E, 9, U, device=dev)
gw0 = torch.empty(E, nir * U, device=dev); gY = torch.zeros(E, Dd, device=dev); gg = torch.empty(N, Dd, U, device=dev)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


bwd = lambda: _lib.tp_bwd(dt, lmax, N, E, U, 9, 9, tab, cgw, csr.row_ptr, csr.ctr, gam, None, Y, w0, go, None, gw0, gY, gg)
names = {0: "full (unroll 1, default)", 256: "full, unroll 2", 1: "no edge arithmetic", 2: "no gY reduction", 32: "no per-centre work", 64: "compute only (no bulk copies)",
         1 | 32: "pipeline only", 2 | 32: "no gY, no centre work", 16: "no Y copies"}
for cps in (0, 3, 2, 1):
    _lib.set_option("tp_stream_cps", cps)
    for dbg, nm in names.items():
        if cps and dbg not in (0, 1 | 32, 64):
            continue
        _lib.set_option("tp_stream3_debug", dbg)
        t = timeit(bwd)
        print(f"cps {cps or 4}  debug {dbg:2d} {nm:32s}: {t:6.1f} us  ({2032 * E / t / 1e3:5.0f} GB/s)", flush=True)
_lib.set_option("tp_stream3_debug", 0)
_lib.set_option("tp_stream_cps", 0)
