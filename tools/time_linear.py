"""Micro-benchmark of ab2_linear on the GPU (tensor-core vs CUDA-core path, stage knock-outs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from allegro_b200 import _lib

M = 461154
dev = "cuda"
shapes = [([64], [96, 64, 96]), ([64], [64, 96]), ([64, 32], [64]), ([64, 64, 64], [64]), ([64], [64]), ([96, 64, 96], [64])]


def run(awid, owid, dtype, debug=0, tc=True, epi=0, accum=False, reps=10):
    K, N = sum(awid), sum(owid)
    a = [torch.randn(M, w, device=dev, dtype=dtype) for w in awid]
    W = torch.randn(K, N, device=dev, dtype=dtype) * 0.1
    o = [torch.zeros(M, w, device=dev, dtype=dtype) for w in owid]
    aux = torch.randn(M, N, device=dev, dtype=dtype) if epi else None
    pk = _lib.linear_pack(W) if tc else None
    _lib.set_option("tc_debug", debug)
    acc = [accum] * len(owid)
    for _ in range(3):
        _lib.linear(a, W, o, o_accum=acc, epi=epi, aux=aux, W_packed=pk)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        _lib.linear(a, W, o, o_accum=acc, epi=epi, aux=aux, W_packed=pk)
    t1.record()
    torch.cuda.synchronize()
    _lib.set_option("tc_debug", 0)
    ms = t0.elapsed_time(t1) / reps
    esz = 4 if dtype == torch.float32 else 2
    byts = M * (K + N * (2 if accum else 1) + (N if epi else 0)) * esz
    return ms, byts / ms / 1e6


shapes += [([128], [192, 128, 192]), ([128, 64], [128]), ([128], [128, 192]), ([192, 128, 192], [128])]  # c3-sized layers
for tma in (1, 0):
    _lib.set_option("linear_tma", tma)
    print(f"--- linear_tma={tma} ({'TMA producer + converter groups' if tma else 'round-1 cp.async producers'}) ---", flush=True)
    for awid, owid in shapes:
        K, N = sum(awid), sum(owid)
        line = f"K={K:3d} N={N:3d}:"
        for name, kw in [("full", {}), ("dsilu", dict(epi=1)), ("accum", dict(accum=True)), ("dsilu+acc", dict(epi=1, accum=True))]:
            ms, gbs = run(awid, owid, torch.float32, **kw)
            line += f"  {name} {ms*1e3:6.0f}us ({gbs:5.0f}GB/s)"
        print(line, flush=True)
_lib.set_option("linear_tma", 1)
