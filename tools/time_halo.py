"""Per-phase timing of the multi-GPU step (run under torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from allegro_b200 import systems
from allegro_b200.halo import SlabDecomposition

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
pos, cell, types = systems.make_positions("c2", (14 * world, 14, 14))
dec = SlabDecomposition(pos, cell, types, 5.0, rank, world)
po = dec.local_positions_from_global(pos)[: dec.n_owned].to(dev)
dec.to(dev)


def timeit(fn, reps=20):
    for _ in range(5):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = time.perf_counter()
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t) / reps * 1e3
    return a.elapsed_time(b) / reps, wall


g = torch.zeros(dec.n_ghost, 3, device=dev, dtype=po.dtype)
go = torch.zeros(dec.n_owned, 3, device=dev, dtype=po.dtype)
e = torch.zeros(1, device=dev, dtype=torch.float64)
res = {
    "fwd_p2p": timeit(lambda: dec.exchange_forward(po)),
    "rev_p2p": timeit(lambda: dec.exchange_reverse(g, go)),
    "allreduce": timeit(lambda: dist.all_reduce(e)),
}
buf = torch.zeros(2048, 3, device=dev, dtype=po.dtype)
out = torch.zeros(world * 2048, 3, device=dev, dtype=po.dtype)
res["all_gather_48KB"] = timeit(lambda: dist.all_gather_into_tensor(out, buf))
if rank == 0:
    for k, (d, w) in res.items():
        print(f"{k:18s} device {d*1e3:8.1f} us   wall {w*1e3:8.1f} us", flush=True)
dist.destroy_process_group()
