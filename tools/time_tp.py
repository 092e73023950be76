"""Micro-benchmark of the tensor-product kernels on c2-shaped data (E=461k, U=32, l_max=2)."""
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
tp1 = Contracter(sh, sh, "1x0e", mul=U)
Y = torch.randn(E, Dd, device=dev)
w0 = torch.randn(E, nir * U, device=dev)
gam = torch.randn(N, Dd, U, device=dev)


def timeit(fn, reps=10):
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


VARIANTS = [  # (label, options)
    ("stream te8", dict(tp_fast=1, tp_stream=1, tp_stream_te=8)),
    ("stream shfl gY", dict(tp_fast=1, tp_stream=1, tp_stream_te=8, tp_stream_gytile=0)),
    ("stream, L1 bwd smem", dict(tp_fast=1, tp_stream=1, tp_stream_te=8, tp_stream3=1, tp_stream_last=0)),
    ("stream3", dict(tp_fast=1, tp_stream=1, tp_stream_te=8, tp_stream3=1)),
    ("stream3 cps3", dict(tp_fast=1, tp_stream=1, tp_stream_te=8, tp_stream3=1, tp_stream_cps=3)),
    ("stream3 cps2", dict(tp_fast=1, tp_stream=1, tp_stream_te=8, tp_stream3=1, tp_stream_cps=2)),
    ("stream3 cps4", dict(tp_fast=1, tp_stream=1, tp_stream_te=8, tp_stream3=1, tp_stream_cps=4)),
    ("stream te8 cps3", dict(tp_fast=1, tp_stream=1, tp_stream_te=8, tp_stream_cps=3)),
    ("stream te8 cps2", dict(tp_fast=1, tp_stream=1, tp_stream_te=8, tp_stream_cps=2)),
    ("stream te16", dict(tp_fast=1, tp_stream=1, tp_stream_te=16)),
    ("r1 smem/split", dict(tp_fast=1, tp_stream=0, tp_variant=1)),
    ("r1 regM", dict(tp_fast=2, tp_stream=0)),
]
ref = {}
for label, opts in VARIANTS:
    for k in ("tp_stream_cps", "tp_stream_te", "tp_variant", "tp_stream3", "tp_stream_gytile", "tp_stream_last"):
        _lib.set_option(k, 1 if k in ("tp_variant", "tp_stream_gytile", "tp_stream_last") else 0)
    for k, v in opts.items():
        _lib.set_option(k, v)
    line = f"{label:16s}:"
    for name, tp, implicit in (("L0", tp0, True), ("L1", tp1, False), ("mid", tp0, False)):
        ijk, _, _ = tp.sparse_table()
        tab, cgw = ijk.to(dev), tp.cgw(dt, dev)
        d_in, d_out = tp.base_dim1, tp.base_dim_out
        g = torch.Generator(device=dev).manual_seed(7)
        Vin = None if implicit else torch.randn(E, d_in, U, device=dev, generator=g)
        Vout = torch.empty(E, d_out, U, device=dev)
        go = torch.randn(E, d_out, U, device=dev, generator=g)
        gVin = None if implicit else torch.empty(E, d_in, U, device=dev)
        gw0 = torch.empty(E, nir * U, device=dev) if implicit else None
        gY = torch.zeros(E, Dd, device=dev) if implicit else None
        gg = torch.empty(N, Dd, U, device=dev)
        fwd = lambda: _lib.tp_fwd(dt, lmax, N, E, U, d_in, d_out, tab, cgw, csr.row_ptr, csr.ctr, gam, Vin, Y, w0 if implicit else None, Vout)
        bwd = lambda: _lib.tp_bwd(dt, lmax, N, E, U, d_in, d_out, tab, cgw, csr.row_ptr, csr.ctr, gam, Vin, Y, w0 if implicit else None, go, gVin, gw0, gY, gg)
        # results of every variant against the first one (all variants compute the same thing)
        if gY is not None:
            gY.zero_()
        fwd(); bwd()
        torch.cuda.synchronize()
        outs = {"Vout": Vout, "gg": gg, "gVin": gVin, "gw0": gw0, "gY": gY}
        for k, v in outs.items():
            if v is None:
                continue
            if (name, k) not in ref:
                ref[(name, k)] = v.clone()
            else:
                err = float((v - ref[(name, k)]).abs().max() / ref[(name, k)].abs().max())
                if err > 1e-4:
                    line += f" [{name}.{k} MISMATCH {err:.1e}]"
        f = timeit(fwd)
        b = timeit(bwd)
        fb = (4 * nir * U + 4 * Dd if implicit else 4 * U * d_in) + 4 + 4 * U * d_out
        bb = fb + (4 * nir * U + 8 * Dd if implicit else 4 * U * d_in)  # inputs + gVout (same size as Vout) + input gradients
        line += f"  {name}: fwd {f:5.0f}us ({fb*E/f/1e3:5.0f}GB/s) bwd {b:5.0f}us ({bb*E/b/1e3:5.0f}GB/s)"
    print(line, flush=True)
