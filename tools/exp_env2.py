"""env_sum unroll and env_bwd_stream CTAs/SM variants at the c2 shapes (E = 461k, U = 32, l_max = 2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from allegro_b200 import _lib, data as D

dev = "cuda"
N, deg, U, lmax = 10976, 42, 32, 2
E, Dd, nir = N * deg, 9, 3
ctr = torch.arange(N).repeat_interleave(deg)
csr = D.build_csr(torch.stack([ctr, (ctr + 1) % N]).to(dev), N)
dt = torch.float32
torch.manual_seed(0)
Y = torch.randn(E, Dd, device=dev); om = torch.randn(E, nir * U, device=dev); gg = torch.randn(N, Dd, U, device=dev)
gom = torch.empty(E, nir * U, device=dev); gY = torch.zeros(E, Dd, device=dev)


def timeit(fn, reps=30):
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


ref = None
for unr in (2, 4):
    _lib.set_option("env_unroll", unr)
    out = _lib.env_sum(dt, lmax, N, U, csr.row_ptr, Y, om, 0.15)
    if ref is None:
        ref = out.clone()
    err = float((out - ref).abs().max() / ref.abs().max())
    print(f"env_sum unroll {unr}: {timeit(lambda: _lib.env_sum(dt, lmax, N, U, csr.row_ptr, Y, om, 0.15)):6.1f} us  (420 B/edge -> {420 * E / timeit(lambda: _lib.env_sum(dt, lmax, N, U, csr.row_ptr, Y, om, 0.15)) / 1e3:5.0f} GB/s)  err {err:.1e}", flush=True)
_lib.set_option("env_unroll", 2)
ref = None
for cps in (0, 8, 6, 5, 4, 3, 2):
    _lib.set_option("env_stream_cps", cps)
    gY.zero_()
    _lib.env_bwd(dt, lmax, U, csr.ctr, Y, om, gg, 0.15, gom, gY, row_ptr=csr.row_ptr)
    torch.cuda.synchronize()
    if ref is None:
        ref = (gom.clone(), gY.clone())
    err = max(float((gom - ref[0]).abs().max() / ref[0].abs().max()), float((gY - ref[1]).abs().max() / ref[1].abs().max()))
    t = timeit(lambda: _lib.env_bwd(dt, lmax, U, csr.ctr, Y, om, gg, 0.15, gom, gY, row_ptr=csr.row_ptr))
    print(f"env_bwd_stream cps {cps}: {t:6.1f} us  (884 B/edge -> {884 * E / t / 1e3:5.0f} GB/s)  err {err:.1e}", flush=True)
_lib.set_option("env_stream_cps", 0)
