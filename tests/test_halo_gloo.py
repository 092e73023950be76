"""world_size-2/3 gloo tests (CPU) of the spatial decomposition + halo exchange host logic:
slab-decomposed energy/forces == the single-process periodic evaluation.  The energy model
is the CPU oracle (test infrastructure stand-in for the CUDA FusedAllegroEnergy)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from allegro_b200 import data as D
from allegro_b200 import systems


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, reps, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        from allegro_b200.halo import DistributedAllegro, SlabDecomposition
        from oracle.model_ref import AllegroEnergyOracle

        pos, cell, types = systems.make_positions(name, reps)
        cfg = systems.CONFIGS[name]
        kw = systems.model_kwargs(name, 42.0, "float64")
        kw.update(num_scalar_features=8, num_tensor_features=4, radial_chemical_embed_dim=8, scalar_embed_mlp_hidden_layers_width=8,
                  allegro_mlp_hidden_layers_width=8, readout_mlp_hidden_layers_width=8, per_type_energy_shifts=[-0.7] * len(cfg["type_names"]))
        model = AllegroEnergyOracle(**kw)
        dec = SlabDecomposition(pos, cell, types, cfg["r_max"], rank, world)
        runner = DistributedAllegro(model, dec)
        pos_owned = dec.local_positions_from_global(pos)[: dec.n_owned]
        e_tot, f_owned, e_atoms = runner(pos_owned)
        # plain numpy payloads: torch tensors cross the queue as shared-memory handles that can die with this process
        q.put((rank, dec.owned.numpy().copy(), f_owned.detach().numpy().copy(), e_atoms.detach().numpy().copy(), float(e_tot), dec.n_ghost))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,name,reps", [(2, "c2", (6, 3, 3)), (3, "c2", (6, 3, 3)), (2, "c5", (6, 3, 3))])
def test_slab_decomposition_matches_periodic_reference(world, name, reps):
    from oracle.model_ref import AllegroOracle

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, reps, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference on the periodic frame
    d = systems.make_system(name, reps)
    cfg = systems.CONFIGS[name]
    kw = systems.model_kwargs(name, 42.0, "float64")
    kw.update(num_scalar_features=8, num_tensor_features=4, radial_chemical_embed_dim=8, scalar_embed_mlp_hidden_layers_width=8,
              allegro_mlp_hidden_layers_width=8, readout_mlp_hidden_layers_width=8, per_type_energy_shifts=[-0.7] * len(cfg["type_names"]))
    ref = AllegroOracle(**kw)(d)
    n = d[D.POSITIONS_KEY].shape[0]
    F = torch.zeros(n, 3, dtype=torch.float64)
    Ea = torch.zeros(n, 1, dtype=torch.float64)
    seen = torch.zeros(n, dtype=torch.long)
    for rank, owned, f, ea, e_tot, n_ghost in res:
        owned = torch.from_numpy(owned)
        F[owned] = torch.from_numpy(f)
        Ea[owned] = torch.from_numpy(ea)
        seen[owned] += 1
        assert n_ghost > 0
        assert e_tot == pytest.approx(ref[D.TOTAL_ENERGY_KEY].item(), rel=1e-11)
    assert bool((seen == 1).all())  # every atom owned exactly once
    assert (Ea - ref[D.PER_ATOM_ENERGY_KEY]).abs().max() < 1e-10
    assert (F - ref[D.FORCE_KEY]).abs().max() < 1e-10
