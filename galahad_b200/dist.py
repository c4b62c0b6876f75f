"""Tree-parallel SLS factor/solve across the GPUs of one node, one process per GPU.

The assembly tree is cut into leaf subtrees (one owner rank each) and a root part (rank 0), as
SSIDS's find_subtree_partition does for its NUMA regions / GPUs (src/ssids/anal.f90:284-459).
There is exactly one exchange step per phase and it is point to point: the packed boundary
fronts (contribution block + delayed columns, the wire format of contrib_type,
src/ssids/contrib.f90:19-33) travel to the root owner with torch.distributed send/recv -- NCCL
over NVLink on GPUs; the CPU tests drive the same schedule over gloo with a mock engine.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


def exchange_to_root(items, rank, nbytes_fn, pack_fn, unpack_fn, device, root=0):
    """items: list of (key, owner).  Owners != root send key's packed bytes to root; root unpacks.
    nbytes_fn(key) -> int (owner side), pack_fn(key, uint8 tensor), unpack_fn(key, uint8 tensor)."""
    moved = 0
    for key, owner in items:
        if owner == root:
            continue
        if rank == owner:
            nb = int(nbytes_fn(key))
            dist.send(torch.tensor([nb], dtype=torch.int64, device=device), dst=root)
            buf = torch.empty(max(nb, 1), dtype=torch.uint8, device=device)
            pack_fn(key, buf)
            dist.send(buf, dst=root)
            moved += nb
        elif rank == root:
            sz = torch.zeros(1, dtype=torch.int64, device=device)
            dist.recv(sz, src=owner)
            nb = int(sz.item())
            buf = torch.empty(max(nb, 1), dtype=torch.uint8, device=device)
            dist.recv(buf, src=owner)
            torch.cuda.synchronize(device) if buf.is_cuda else None   # NCCL is async; the engine reads on its own stream
            unpack_fn(key, buf, nb)
            moved += nb
    return moved


class DistSolver:
    """analyse -> factor -> solve over the default process group.  Every rank passes the same
    matrix, order and right-hand side (the callers' data is replicated host data)."""

    def __init__(self, device=None, **opts):
        from ._lib import Solver
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.dev_index = torch.cuda.current_device() if device is None else device
        self.device = torch.device("cuda", self.dev_index)
        self.s = Solver(device=self.dev_index, **opts)
        self.boundaries = []
        self.bytes_moved = 0

    def analyse(self, n, ptr, row, order):
        flag = self.s.analyse(n, ptr, row, order)
        self.n = n
        s = self.s
        f = s.L.b200ldlt_dist_init(s.akeep, C.byref(s.fkeep), self.rank, self.world, C.byref(s.opt), C.byref(s.inform))
        assert f >= 0, f"dist_init flag {f}"
        nb = s.L.b200ldlt_dist_nboundary(s.akeep)
        self.boundaries = []
        for i in range(nb):
            node, owner = C.c_int(), C.c_int()
            s.L.b200ldlt_dist_boundary(s.akeep, i, C.byref(node), C.byref(owner))
            self.boundaries.append((node.value, owner.value))
        return flag

    def factor(self, val, posdef=False):
        s = self.s
        val = np.ascontiguousarray(val, np.float64)
        vp = val.ctypes.data_as(C.POINTER(C.c_double))
        f1 = s.L.b200ldlt_dist_factor(s.akeep, s.fkeep, 1, int(posdef), vp, C.byref(s.opt), C.byref(s.inform))
        t1 = s.inform.time_factor
        stats = np.array([s.inform.num_neg, s.inform.num_two, s.inform.num_delay, 0, min(f1, 0)], dtype=np.int64)
        self.bytes_moved = exchange_to_root(
            self.boundaries, self.rank,
            lambda node: s.L.b200ldlt_dist_contrib_bytes(s.fkeep, node),
            lambda node, buf: s.L.b200ldlt_dist_contrib_pack(s.fkeep, node, C.c_void_p(buf.data_ptr())),
            lambda node, buf, nb: s.L.b200ldlt_dist_contrib_unpack(s.fkeep, node, C.c_void_p(buf.data_ptr()), nb),
            self.device)
        t2 = 0.0
        if self.rank == 0:
            f2 = s.L.b200ldlt_dist_factor(s.akeep, s.fkeep, 2, int(posdef), vp, C.byref(s.opt), C.byref(s.inform))
            t2 = s.inform.time_factor
            # after phase 2 the counters cover every front factorised on rank 0 (leaves + root part)
            stats = np.array([s.inform.num_neg, s.inform.num_two, s.inform.num_delay, 0, min(f1, f2, 0)], dtype=np.int64)
        self.t_leaves, self.t_root = t1, t2
        t = torch.from_numpy(stats).to(self.device)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        tot = t.cpu().numpy()
        self.num_neg, self.num_two, self.num_delay = int(tot[0]), int(tot[1]), int(tot[2])
        return int(tot[4])

    def solve(self, b):
        s = self.s
        x = torch.from_numpy(np.ascontiguousarray(b, np.float64)).to(self.device)
        torch.cuda.synchronize(self.device)
        f = s.L.b200ldlt_dist_solve(s.fkeep, 1, C.c_void_p(x.data_ptr()), None)
        assert f >= 0, f
        exchange_to_root(
            self.boundaries, self.rank,
            lambda node: s.L.b200ldlt_dist_xtail_bytes(s.fkeep, node),
            lambda node, buf: s.L.b200ldlt_dist_xtail_pack(s.fkeep, node, C.c_void_p(buf.data_ptr())),
            lambda node, buf, nb: s.L.b200ldlt_dist_xtail_unpack(s.fkeep, node, C.c_void_p(buf.data_ptr())),
            self.device)
        xp = torch.zeros(self.n, dtype=torch.float64, device=self.device)
        if self.rank == 0:
            f = s.L.b200ldlt_dist_solve(s.fkeep, 2, C.c_void_p(x.data_ptr()), None)
            assert f >= 0, f
            torch.cuda.synchronize(self.device)
            s.L.b200ldlt_dist_xp_get(s.fkeep, C.c_void_p(xp.data_ptr()), self.n)
        if self.world > 1:
            dist.broadcast(xp, src=0)
            torch.cuda.synchronize(self.device)
        src = None if self.rank == 0 else C.c_void_p(xp.data_ptr())
        f = s.L.b200ldlt_dist_solve(s.fkeep, 3, C.c_void_p(x.data_ptr()), src)
        assert f >= 0, f
        f = s.L.b200ldlt_dist_solve(s.fkeep, 4, C.c_void_p(x.data_ptr()), None)
        assert f >= 0, f
        if self.world > 1:
            dist.all_reduce(x, op=dist.ReduceOp.SUM)
        return x.cpu().numpy()
