"""Multi-GPU parity (needs >= 2 CUDA devices, skipped otherwise): slab decomposition + NCCL halo with the
CUDA model on every rank == the single-GPU periodic evaluation of the same frame."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, mode="nccl", reps=(6, 3, 3)):
    import torch.distributed as dist

    from allegro_b200 import systems
    from allegro_b200.halo import DistributedAllegro, SlabDecomposition
    from allegro_b200.model import AllegroModel

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        pos, cell, types = systems.make_positions("c2", reps)
        kw = systems.model_kwargs("c2", 42.0, "float64")
        model = AllegroModel(**kw).to(dev).model
        # "p2p": NVLink peer-memory halo (no NCCL per step) + CUDA neighbour list straight into CSR
        dec = SlabDecomposition(pos, cell, types, 5.0, rank, world, device=dev if mode == "p2p" else None)
        pos_owned = dec.local_positions_from_global(pos)[: dec.n_owned].to(dev)
        dec.to(dev)
        p2p = None
        if mode == "p2p":
            from allegro_b200.halo import P2PHalo

            assert dec.csr is not None
            p2p = P2PHalo(dec, dev)
        runner = DistributedAllegro(model, dec, p2p=p2p)
        e_tot, f_owned, e_atoms = runner(pos_owned)
        # the same step replayed from a CUDA graph (NCCL halo captured) on displaced positions
        from allegro_b200.halo import GraphedDistributedAllegro

        g = torch.Generator().manual_seed(7 + rank)
        pos2 = pos_owned + 0.05 * torch.randn(pos_owned.shape, generator=g, dtype=pos_owned.dtype).to(dev)
        e2, f2, _ = runner(pos2)
        graphed = GraphedDistributedAllegro(runner, pos_owned)
        eg, fg, _ = graphed(pos2)
        torch.cuda.synchronize()
        gerr = max(float((fg - f2).abs().max()), abs(float(eg) - float(e2)))
        if p2p is not None:
            # a few more replays: the mailbox parity / step-number protocol over several steps
            for _ in range(5):
                eg, fg, _ = graphed(pos2)
            torch.cuda.synchronize()
            gerr = max(gerr, float((fg - f2).abs().max()), abs(float(eg) - float(e2)), float(p2p.error()))
        # plain numpy payloads: torch tensors travel as shared-memory handles that die with this process
        q.put((rank, dec.owned.cpu().numpy(), f_owned.cpu().numpy(), e_atoms.cpu().numpy(), float(e_tot), gerr))
        # a live graph with captured NCCL kernels blocks communicator teardown: flush the queue and
        # leave without the destructors (the parent checks exit code 0)
        dist.barrier()
        torch.cuda.synchronize()
        q.close()
        q.join_thread()
        os._exit(0)
    except BaseException:
        dist.destroy_process_group()
        raise


@pytest.mark.parametrize("mode,reps", [("nccl", (6, 3, 3)), ("p2p", (8, 5, 5))])
def test_halo_matches_single_gpu(mode, reps):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from allegro_b200 import data as D
    from allegro_b200 import systems
    from allegro_b200.model import AllegroModel

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, mode, reps)) for r in range(world)]
    for p in procs:
        p.start()
    import queue
    import time

    res, deadline = [], time.time() + 300
    while len(res) < world:  # fail fast when a worker dies instead of waiting out the queue timeout
        try:
            res.append(q.get(timeout=2))
        except queue.Empty:
            dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
            assert not dead, f"worker exited with {dead}"
            assert time.time() < deadline, "timed out waiting for the workers"
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d = systems.make_system("c2", reps)
    kw = systems.model_kwargs("c2", 42.0, "float64")
    ref = AllegroModel(**kw).to("cuda:0")({k: v.to("cuda:0") for k, v in d.items()})
    n = d[D.POSITIONS_KEY].shape[0]
    F = torch.zeros(n, 3, dtype=torch.float64)
    Ea = torch.zeros(n, 1, dtype=torch.float64)
    for rank, owned, f, ea, e_tot, gerr in res:
        assert gerr < 1e-9
        owned = torch.from_numpy(owned)
        F[owned] = torch.from_numpy(f).double()
        Ea[owned] = torch.from_numpy(ea).double()
        assert e_tot == pytest.approx(ref[D.TOTAL_ENERGY_KEY].item(), rel=1e-10)
    assert (Ea - ref[D.PER_ATOM_ENERGY_KEY].cpu()).abs().max() < 1e-9
    assert (F - ref[D.FORCE_KEY].cpu()).abs().max() < 1e-9
