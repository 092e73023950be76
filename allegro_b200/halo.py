"""Spatial domain decomposition with a single-r_max halo.

Allegro is strictly local: E_i depends only on atoms within r_max of i
(/root/reference/tests/model/test_allegro.py:68-70; the environment sum only runs over edges
centred on i, allegro/nn/_strided/_contract.py:199-205).  So atoms are split into 1-D slabs
along x, one per rank / GPU; each rank evaluates the edges of its OWNED centres, with
neighbours taken from owned + ghost atoms in exactly the reference's ghost-atom data format
(ghosts appended after the locals, neighbour index >= N_local,
/root/reference/allegro/_compile.py:41-61).

Per force evaluation the only communication is
  1. forward halo : positions of boundary atoms -> neighbouring slabs' ghosts
  2. reverse halo : dE/dpos accumulated on ghosts -> added on the owning rank
  3. one all-reduce of the scalar total energy
through torch.distributed point-to-point ops (NCCL over NVLink on the GPU box, gloo in the
CPU tests).  There is no reference counterpart (the reference delegates this to LAMMPS' MPI).

The decomposition itself is set up from the global synthetic frame, which every rank can
generate deterministically, so set-up needs no communication.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import data as D


class SlabDecomposition:
    """Static exchange plan of one rank (valid as long as the neighbour list is)."""

    def __init__(self, pos: torch.Tensor, cell: torch.Tensor, types: torch.Tensor, r_max: float, rank: int, world: int, device=None):
        assert world >= 2, "use the plain periodic path on one rank"
        cell = cell.view(3, 3)
        assert bool((cell - torch.diag(torch.diagonal(cell))).abs().max() == 0), "slab decomposition needs an orthorhombic box"
        Lx = float(cell[0, 0])
        width = Lx / world
        assert width >= r_max, "slab thinner than r_max: neighbours beyond the adjacent slab"
        self.rank, self.world, self.r_max, self.Lx = rank, world, float(r_max), Lx
        x = pos[:, 0] - torch.floor(pos[:, 0] / Lx) * Lx
        slab = torch.clamp((x / width).long(), max=world - 1)

        def owned_of(r):
            return torch.nonzero(slab == r).squeeze(-1)

        def send_lists(r):
            ids = owned_of(r)
            xr = x[ids]
            lo, hi = r * width, (r + 1) * width
            return ids[xr >= hi - r_max], ids[xr < lo + r_max]  # (to the right neighbour, to the left neighbour)

        self.left, self.right = (rank - 1) % world, (rank + 1) % world
        self.owned = owned_of(rank)
        self.n_owned = int(self.owned.shape[0])
        my_to_right, my_to_left = send_lists(rank)
        # positions in the local owned array
        inv = torch.full((pos.shape[0],), -1, dtype=torch.long)
        inv[self.owned] = torch.arange(self.n_owned)
        self.send_right_idx = inv[my_to_right]
        self.send_left_idx = inv[my_to_left]
        # what arrives: left neighbour's "to the right" list, right neighbour's "to the left" list
        left_ids = send_lists(self.left)[0]
        right_ids = send_lists(self.right)[1]
        self.n_ghost_left, self.n_ghost_right = int(left_ids.shape[0]), int(right_ids.shape[0])
        self.ghost_global = torch.cat([left_ids, right_ids])
        self.n_ghost = self.n_ghost_left + self.n_ghost_right
        # periodic wrap of the ghost images along x
        self.shift_left = -Lx if rank == 0 else 0.0
        self.shift_right = Lx if rank == world - 1 else 0.0
        self.types_local = torch.cat([types[self.owned], types[self.ghost_global]])
        self.global_ids_local = torch.cat([self.owned, self.ghost_global])
        # local periodic cell: x is open (ghosts cover it), y and z stay periodic
        self.cell = cell.clone()
        self.pbc = (False, True, True)
        # unwrapped x so that slab + ghosts form one contiguous block along x
        pos_local = self.local_positions_from_global(pos)
        self.csr, self.shift_vec = None, None
        self.edge_index = self.edge_cell_shift = None
        dev = torch.device(device) if device is not None else None
        if dev is not None and dev.type == "cuda" and D.csr_supported(pos_local.to(dev), r_max, self.cell, self.pbc):
            # CUDA cell list over owned + ghost atoms, rows for the owned centres only, straight into the kernels' CSR:
            # the int64 COO list of the local frame (6.5M edges per rank for the 1M-atom box) is never built
            self.csr, self.shift_vec = D.neighbor_csr(pos_local.to(dev), r_max, self.cell.to(dev), self.pbc, n_centres=self.n_owned)
            self.n_edges = self.csr.num_edges
        else:
            ei, sh = D.neighbor_list(pos_local, r_max, self.cell, self.pbc)
            keep = ei[0] < self.n_owned  # edges of owned centres only
            self.edge_index = ei[:, keep].contiguous()
            self.edge_cell_shift = sh[keep].contiguous()
            self.n_edges = int(self.edge_index.shape[1])

    # positions of owned + ghost atoms taken from a global frame (set-up / tests)
    def local_positions_from_global(self, pos: torch.Tensor) -> torch.Tensor:
        Lx = self.Lx
        p = pos.clone()
        p[:, 0] = p[:, 0] - torch.floor(p[:, 0] / Lx) * Lx
        gl = p[self.ghost_global[: self.n_ghost_left]].clone()
        gr = p[self.ghost_global[self.n_ghost_left :]].clone()
        gl[:, 0] += self.shift_left
        gr[:, 0] += self.shift_right
        return torch.cat([p[self.owned], gl, gr], 0)

    def to(self, device) -> "SlabDecomposition":
        for k in ("owned", "send_right_idx", "send_left_idx", "ghost_global", "types_local", "global_ids_local", "cell",
                  "edge_index", "edge_cell_shift"):
            if getattr(self, k) is not None:
                setattr(self, k, getattr(self, k).to(device))
        return self

    # ---- communication --------------------------------------------------------------------
    def _peers(self):
        """-> list of (peer, [send segments], [recv segment sizes]) in a canonical order that
        both sides agree on (right-going list first, then left-going)."""
        if self.left == self.right:  # world == 2: both neighbours are the same rank
            return [(self.left, ["right", "left"], ["from_left", "from_right"])]
        return [(self.right, ["right"], ["from_right"]), (self.left, ["left"], ["from_left"])]

    def exchange_forward(self, pos_owned: torch.Tensor) -> torch.Tensor:
        """positions of owned atoms [n_owned,3] -> ghost positions [n_ghost,3] (wrap applied)."""
        seg = {"right": pos_owned.index_select(0, self.send_right_idx), "left": pos_owned.index_select(0, self.send_left_idx)}
        n_in = {"from_left": self.n_ghost_left, "from_right": self.n_ghost_right}
        ops, bufs = [], []
        for peer, sends, recvs in self._peers():
            sbuf = torch.cat([seg[s] for s in sends], 0).contiguous()
            rbuf = torch.empty(sum(n_in[r] for r in recvs), 3, dtype=pos_owned.dtype, device=pos_owned.device)
            ops += [dist.P2POp(dist.isend, sbuf, peer), dist.P2POp(dist.irecv, rbuf, peer)]
            bufs.append((recvs, rbuf, sbuf))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        got = {}
        for recvs, rbuf, _ in bufs:
            o = 0
            for r in recvs:
                got[r] = rbuf[o : o + n_in[r]]
                o += n_in[r]
        gl, gr = got["from_left"].clone(), got["from_right"].clone()
        gl[:, 0] += self.shift_left
        gr[:, 0] += self.shift_right
        return torch.cat([gl, gr], 0)

    def exchange_reverse(self, g_ghost: torch.Tensor, g_owned: torch.Tensor) -> torch.Tensor:
        """add the gradient accumulated on ghosts [n_ghost,3] onto the owners; returns g_owned."""
        seg = {"from_left": g_ghost[: self.n_ghost_left], "from_right": g_ghost[self.n_ghost_left :]}
        n_out = {"right": int(self.send_right_idx.shape[0]), "left": int(self.send_left_idx.shape[0])}
        ops, bufs = [], []
        for peer, sends, recvs in self._peers():
            # mirror of the forward plan: what I received from `peer` goes back to it
            sbuf = torch.cat([seg[r] for r in recvs], 0).contiguous()
            rbuf = torch.empty(sum(n_out[s] for s in sends), 3, dtype=g_ghost.dtype, device=g_ghost.device)
            ops += [dist.P2POp(dist.isend, sbuf, peer), dist.P2POp(dist.irecv, rbuf, peer)]
            bufs.append((sends, rbuf, sbuf))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        idx = {"right": self.send_right_idx, "left": self.send_left_idx}
        for sends, rbuf, _ in bufs:
            o = 0
            for s in sends:
                g_owned.index_add_(0, idx[s], rbuf[o : o + n_out[s]])
                o += n_out[s]
        return g_owned

    def halo_bytes_per_step(self, itemsize: int) -> int:
        n = int(self.send_right_idx.shape[0]) + int(self.send_left_idx.shape[0]) + self.n_ghost
        return n * 3 * itemsize


class P2PHalo:
    """The three per-step exchanges over NVLink peer memory (csrc/halo_p2p.cu): every rank owns a mailbox that its peers
    map through CUDA IPC; senders store rows straight into the receiver's mailbox and publish the step number, receivers
    spin on it.  Kernels only -- the step stays one CUDA graph and makes no NCCL call.  Set-up (once) uses
    torch.distributed to exchange the IPC handles."""

    def __init__(self, dec: SlabDecomposition, device):
        import ctypes as C

        from . import _lib

        self._lib, self._C = _lib, C
        self.dec, self.dev = dec, torch.device(device)
        lib = _lib.load()
        world, rank = dec.world, dec.rank
        mx = torch.tensor([max(int(dec.send_right_idx.shape[0]), int(dec.send_left_idx.shape[0]), dec.n_ghost_left, dec.n_ghost_right, 1)],
                          device=self.dev)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        self.max_rows = int(mx)
        nbytes = int(lib.ab2_p2p_mailbox_bytes(self.max_rows, world))
        ptr = C.c_void_p()
        _lib._check(lib.ab2_p2p_alloc(nbytes, C.byref(ptr)))
        self.mailbox = ptr.value
        hbuf = C.create_string_buffer(64)
        _lib._check(lib.ab2_p2p_get_handle(C.c_void_p(self.mailbox), hbuf))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(hbuf.raw))
        self.peers = []
        for r, h in enumerate(handles):
            if r == rank:
                self.peers.append(self.mailbox)
            else:
                q = C.c_void_p()
                _lib._check(lib.ab2_p2p_open_handle(C.create_string_buffer(h, 64), C.byref(q)))
                self.peers.append(q.value)
        self.peers_dev = torch.tensor(self.peers, dtype=torch.int64, device=self.dev)
        self.step = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.done = torch.zeros(4, dtype=torch.int32, device=self.dev)
        self.e_tot = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self.pos_local = None
        self.send_right = dec.send_right_idx.to(self.dev).contiguous()
        self.send_left = dec.send_left_idx.to(self.dev).contiguous()
        # periodic wrap along x as the RECEIVER sees my atoms (rank 0's left ghosts sit at x - Lx, ...)
        self.shift_to_right = -dec.Lx if rank == world - 1 else 0.0
        self.shift_to_left = dec.Lx if rank == 0 else 0.0
        torch.cuda.synchronize(self.dev)
        dist.barrier()

    def _ptr(self, t, row_offset=0):
        return self._C.c_void_p(t.data_ptr() + row_offset * 3 * t.element_size())

    def forward(self, pos_owned: torch.Tensor) -> torch.Tensor:
        """positions of my atoms -> [owned | ghosts from the left | ghosts from the right] (static buffer)."""
        lib, dec, C = self._lib.load(), self.dec, self._C
        st = self._lib._stream()
        dt = self._lib.DTYPE_ENUM[pos_owned.dtype]
        if self.pos_local is None or self.pos_local.dtype != pos_owned.dtype:
            self.pos_local = torch.zeros(dec.n_owned + dec.n_ghost, 3, dtype=pos_owned.dtype, device=self.dev)
        pos_owned = pos_owned.contiguous()
        self.pos_local[: dec.n_owned].copy_(pos_owned)
        ck = self._lib._check
        ck(lib.ab2_p2p_begin(C.c_void_p(self.step.data_ptr()), st))
        # I am the LEFT neighbour of my right peer (its slot 0) and the RIGHT neighbour of my left peer (its slot 1)
        ck(lib.ab2_p2p_push_rows(dt, 0, 0, self._ptr(pos_owned), C.c_void_p(self.send_right.data_ptr()), int(self.send_right.shape[0]), self.shift_to_right,
                                 C.c_void_p(self.peers[dec.right]), self.max_rows, dec.world, C.c_void_p(self.step.data_ptr()),
                                 C.c_void_p(self.done.data_ptr()), st))
        ck(lib.ab2_p2p_push_rows(dt, 0, 1, self._ptr(pos_owned), C.c_void_p(self.send_left.data_ptr()), int(self.send_left.shape[0]), self.shift_to_left,
                                 C.c_void_p(self.peers[dec.left]), self.max_rows, dec.world, C.c_void_p(self.step.data_ptr()),
                                 C.c_void_p(self.done.data_ptr() + 4), st))
        ck(lib.ab2_p2p_wait_unpack(dt, 0, 0, C.c_void_p(self.mailbox), self.max_rows, dec.world, C.c_void_p(self.step.data_ptr()), dec.n_ghost_left,
                                   self._ptr(self.pos_local, dec.n_owned), None, 0, st))
        ck(lib.ab2_p2p_wait_unpack(dt, 0, 1, C.c_void_p(self.mailbox), self.max_rows, dec.world, C.c_void_p(self.step.data_ptr()), dec.n_ghost_right,
                                   self._ptr(self.pos_local, dec.n_owned + dec.n_ghost_left), None, 0, st))
        self._lib.PROF.launches += 5
        return self.pos_local

    def reverse(self, g_local: torch.Tensor, g_owned: torch.Tensor) -> torch.Tensor:
        """gradients accumulated on my ghosts go back to their owners; what my neighbours accumulated on my boundary atoms
        is added into g_owned (left neighbour's rows first, then the right one's: fixed order)."""
        lib, dec, C = self._lib.load(), self.dec, self._C
        st = self._lib._stream()
        dt = self._lib.DTYPE_ENUM[g_local.dtype]
        g_local = g_local.contiguous()
        ck = self._lib._check
        ck(lib.ab2_p2p_push_rows(dt, 1, 1, self._ptr(g_local, dec.n_owned), None, dec.n_ghost_left, 0.0, C.c_void_p(self.peers[dec.left]), self.max_rows,
                                 dec.world, C.c_void_p(self.step.data_ptr()), C.c_void_p(self.done.data_ptr() + 8), st))
        ck(lib.ab2_p2p_push_rows(dt, 1, 0, self._ptr(g_local, dec.n_owned + dec.n_ghost_left), None, dec.n_ghost_right, 0.0,
                                 C.c_void_p(self.peers[dec.right]), self.max_rows, dec.world, C.c_void_p(self.step.data_ptr()),
                                 C.c_void_p(self.done.data_ptr() + 12), st))
        dto = self._lib.DTYPE_ENUM[g_owned.dtype]
        ck(lib.ab2_p2p_wait_unpack(dto, 1, 0, C.c_void_p(self.mailbox), self.max_rows, dec.world, C.c_void_p(self.step.data_ptr()),
                                   int(self.send_left.shape[0]), self._ptr(g_owned), C.c_void_p(self.send_left.data_ptr()), 1, st))
        ck(lib.ab2_p2p_wait_unpack(dto, 1, 1, C.c_void_p(self.mailbox), self.max_rows, dec.world, C.c_void_p(self.step.data_ptr()),
                                   int(self.send_right.shape[0]), self._ptr(g_owned), C.c_void_p(self.send_right.data_ptr()), 1, st))
        self._lib.PROF.launches += 4
        return g_owned

    def energy(self, e_local: torch.Tensor) -> torch.Tensor:
        """sum of one fp64 scalar over all ranks, in rank order (bitwise identical everywhere)."""
        lib, dec, C = self._lib.load(), self.dec, self._C
        self._e_in = e_local.detach().double().reshape(1).contiguous()
        self._lib._check(lib.ab2_p2p_allreduce_energy(C.c_void_p(self._e_in.data_ptr()), dec.rank, dec.world, self.max_rows,
                                                      C.c_void_p(self.peers_dev.data_ptr()), C.c_void_p(self.mailbox), C.c_void_p(self.step.data_ptr()),
                                                      C.c_void_p(self.e_tot.data_ptr()), self._lib._stream()))
        self._lib.PROF.launches += 2
        return self.e_tot

    def error(self) -> int:
        return int(self._lib.load().ab2_p2p_error(self._C.c_void_p(self.mailbox), self.max_rows, self.dec.world, self._lib._stream()))


class DistributedAllegro:
    """Energy + forces of a slab-decomposed frame.  ``energy_model(data) -> data`` must write
    ``atomic_energy`` for the local atoms (owned first) and be differentiable w.r.t. ``pos``
    (FusedAllegroEnergy on the GPU; any stand-in in the CPU tests)."""

    def __init__(self, energy_model: Callable[[D.Type], D.Type], dec: SlabDecomposition, p2p: Optional["P2PHalo"] = None):
        self.model, self.dec, self.p2p = energy_model, dec, p2p

    def __call__(self, pos_owned: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """-> (total energy [all ranks], forces on owned atoms [n_owned,3], owned atomic energies)."""
        dec = self.dec
        if self.p2p is not None:
            pos_local = self.p2p.forward(pos_owned.detach())
        else:
            ghosts = dec.exchange_forward(pos_owned.detach())
            pos_local = torch.cat([pos_owned.detach(), ghosts], 0).requires_grad_(True)
        data = {D.POSITIONS_KEY: pos_local, D.ATOM_TYPE_KEY: dec.types_local, D.CELL_KEY: dec.cell}
        if dec.csr is not None:
            data[D.CSR_KEY], data[D.EDGE_SHIFT_VEC_KEY] = dec.csr, dec.shift_vec
        else:
            data[D.EDGE_INDEX_KEY], data[D.EDGE_CELL_SHIFT_KEY] = dec.edge_index, dec.edge_cell_shift
        if hasattr(self.model, "energy_and_forces"):
            # autograd-free CUDA path: forces on owned AND ghost atoms come out of one pass
            data[D.POSITIONS_KEY] = pos_local.detach()
            out = self.model.energy_and_forces(data)
            e_atoms = out[D.PER_ATOM_ENERGY_KEY][: dec.n_owned]
            e_local = e_atoms.sum()
            g = -out[D.FORCE_KEY]
        else:
            with torch.enable_grad():
                out = self.model(data)
                e_atoms = out[D.PER_ATOM_ENERGY_KEY][: dec.n_owned]
                e_local = e_atoms.sum()
                (g,) = torch.autograd.grad(e_local, pos_local)
        g_owned = g[: dec.n_owned].clone()
        if self.p2p is not None:
            self.p2p.reverse(g, g_owned)
            # the mailbox protocol sums into a persistent buffer: hand out a copy, so that a caller holding the result of one
            # step does not see it change under the next one (under graph capture the copy is the graph's static output)
            e_tot = self.p2p.energy(e_local).clone()
            return e_tot, -g_owned, e_atoms.detach()
        dec.exchange_reverse(g[dec.n_owned :].contiguous(), g_owned)
        e_tot = e_local.detach().double().clone().reshape(1)
        dist.all_reduce(e_tot)
        return e_tot, -g_owned, e_atoms.detach()


class GraphedDistributedAllegro:
    """One slab-decomposed energy+forces step captured into a CUDA graph, NCCL halo included.

    The eager step is ~60 ctypes launches plus three torch.distributed calls; with one Python
    process per GPU sharing the box's host cores the launch path becomes the bottleneck at 4-8
    ranks.  NCCL point-to-point and all-reduce kernels are capturable once their communicators
    exist (warm-up does that), so the whole step -- forward halo, compute, reverse halo,
    energy all-reduce -- replays as one graph launch per rank."""

    def __init__(self, runner: DistributedAllegro, pos_owned: torch.Tensor, warmup: int = 3):
        from . import _lib

        self.runner = runner
        self.static_pos = pos_owned.detach().clone()
        prof = _lib.PROF.enabled
        _lib.PROF.enabled = False
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(warmup):
                runner(self.static_pos)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        dist.barrier()
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.PROF.launches
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = runner(self.static_pos)
        self.launches_per_replay = _lib.PROF.launches - n0
        _lib.PROF.enabled = prof
        self._lib = _lib

    def __call__(self, pos_owned: Optional[torch.Tensor] = None):
        if pos_owned is not None:
            self.static_pos.copy_(pos_owned, non_blocking=True)
        self.graph.replay()
        self._lib.PROF.launches += self.launches_per_replay
        return self.out
