"""CPU emulation of mixed bf16 / fp32 STORAGE on the c2 architecture (host pipeline + tests/kernel_spec.py, arithmetic in fp64,
selected tensors rounded to bf16 when stored): which tensors can be bf16 while energies/forces stay within 1e-3 of the fp64
oracle (BASELINE configs[1] names bf16)?  Test infrastructure / design experiment, not product code."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import kernel_spec
from allegro_b200 import _lib, systems, data as D
from allegro_b200.model import AllegroModel
from allegro_b200.model.allegro_models import FusedAllegroEnergy
from allegro_b200.nn import _pipeline
from oracle.model_ref import AllegroOracle

for name in kernel_spec.ALL:
    setattr(_lib, name, getattr(kernel_spec, name))
FusedAllegroEnergy.core = lambda self: self._core_for(torch.device("cpu"))

PHASE = {"bwd": False}
ROUND = set()

def rb(t):  # round a tensor in place to bf16 precision
    t.copy_(t.to(torch.bfloat16).to(t.dtype))

_lin, _tpf, _tpb, _envb = kernel_spec.linear, kernel_spec.tp_fwd, kernel_spec.tp_bwd, kernel_spec.env_bwd
def linear(a_segs, W, o_segs, **kw):
    _lin(a_segs, W, o_segs, **kw)
    key = "lin_bwd" if PHASE["bwd"] else "lin_fwd"
    if key in ROUND:
        # accumulate outputs: the sum is re-rounded, as a bf16 buffer would
        for o in o_segs:
            rb(o)
def tp_fwd(*a):
    _tpf(*a)
    if "V" in ROUND: rb(a[-1])
def tp_bwd(dtype, lmax, N, E, U, d_in, d_out, tab, cgw, row_ptr, ctr, gamma, Vin, Y, w0, gVout, gVin, gw0, gY, ggamma):
    _tpb(dtype, lmax, N, E, U, d_in, d_out, tab, cgw, row_ptr, ctr, gamma, Vin, Y, w0, gVout, gVin, gw0, gY, ggamma)
    if "gV" in ROUND:
        if gVin is not None: rb(gVin)
        if gw0 is not None: rb(gw0)
def env_bwd(dtype, lmax, U, ctr, Y, w, ggamma, sf, gw, gY, row_ptr=None):
    _envb(dtype, lmax, U, ctr, Y, w, ggamma, sf, gw, gY, row_ptr=row_ptr)
    if "gw" in ROUND: rb(gw)
_lib.linear, _lib.tp_fwd, _lib.tp_bwd, _lib.env_bwd = linear, tp_fwd, tp_bwd, env_bwd
_bw = _pipeline.AllegroCore.backward
def backward(self, sv, gEi):
    PHASE["bwd"] = True
    try:
        return _bw(self, sv, gEi)
    finally:
        PHASE["bwd"] = False
_pipeline.AllegroCore.backward = backward
# the radial MLP backward runs after core.backward: count it as backward too
_ef = _pipeline.energy_forces

d = systems.make_system("c2", 3)
kw = systems.model_kwargs("c2", d[D.EDGE_INDEX_KEY].shape[1] / d[D.POSITIONS_KEY].shape[0], "float64")
oracle = AllegroOracle(**kw)
ref = oracle(d)
model = AllegroModel(**kw)
model.load_state_dict(oracle.state_dict())

def run(tag, rounds):
    ROUND.clear(); ROUND.update(rounds)
    out = model.model._energy_and_forces(dict(d), False)
    ee = float((out[D.PER_ATOM_ENERGY_KEY] - ref[D.PER_ATOM_ENERGY_KEY]).abs().max() / ref[D.PER_ATOM_ENERGY_KEY].abs().max())
    ef = float((out[D.FORCE_KEY] - ref[D.FORCE_KEY]).abs().max() / ref[D.FORCE_KEY].abs().max())
    print(f"{tag:46s} E {ee:.2e}  F {ef:.2e}", flush=True)

run("nothing rounded (fp64 storage)", [])
run("V (tp_fwd outputs)", ["V"])
run("forward linear outputs (X, omega, w0, hidden)", ["lin_fwd"])
run("all forward activations", ["V", "lin_fwd"])
run("gV/gw0 (tp_bwd outputs)", ["gV"])
run("gomega (env_bwd outputs)", ["gw"])
run("backward linear outputs", ["lin_bwd"])
run("all backward tensors", ["gV", "gw", "lin_bwd"])
run("everything", ["V", "lin_fwd", "gV", "gw", "lin_bwd"])
run("V + gV only (the 9x-wide tensors)", ["V", "gV"])
run("V + gV + gomega", ["V", "gV", "gw"])
