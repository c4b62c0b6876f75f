"""Tree-parallel SLS factor/solve across the GPUs of one node, one process per GPU.

The assembly tree is mapped onto the ranks by proportional mapping (csrc/analyse.cpp), the multi-level
generalisation of SSIDS's find_subtree_partition (src/ssids/anal.f90:284-459): leaf subtrees go to single
ranks, the fronts above them to the first rank of the group that shares their subtree, and stages are
numbered so that stage k only needs fronts of earlier stages from other ranks.  After every stage the
packed boundary fronts (contribution block + delayed columns, the wire format of contrib_type,
src/ssids/contrib.f90:19-33) travel point to point to the owner of their parent with torch.distributed
send/recv -- NCCL over NVLink on GPUs; the CPU tests drive the same schedule over gloo with a mock engine.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


_bufs = {}


def _buffer(tag, nbytes, device):
    """A reusable byte buffer per (tag, device): the exchange of every factorisation moves the same fronts."""
    key = (tag, str(device))
    buf = _bufs.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes * 1.1), 1), dtype=torch.uint8, device=device)
        _bufs[key] = buf
    return buf[:max(nbytes, 1)]


def exchange(items, rank, nbytes_fn, pack_fn, unpack_fn, device):
    """items: list of (key, src_rank, dst_rank), the same list in the same order on every rank.  src packs
    key's bytes and sends them to dst, dst unpacks; pairs that do not involve this rank are skipped.
    nbytes_fn(key) -> int (src side), pack_fn(key, uint8 tensor), unpack_fn(key, uint8 tensor, nbytes).
    The sizes of all items travel in one all-reduce (every item has exactly one source), then all transfers of the
    step are posted together (batched point-to-point operations: they share the links instead of queueing behind
    each other).  Returns the bytes this rank sent or received."""
    moving = [(key, src, dst) for key, src, dst in items if src != dst]
    if not moving:
        return 0
    sizes = torch.zeros(len(moving), dtype=torch.int64, device=device)
    for j, (key, src, dst) in enumerate(moving):
        if rank == src:
            sizes[j] = int(nbytes_fn(key))
    dist.all_reduce(sizes, op=dist.ReduceOp.SUM)
    sz = [int(v) for v in sizes.cpu().tolist()]
    ops, incoming, moved = [], [], 0
    for j, (key, src, dst) in enumerate(moving):
        if rank == src:
            buf = _buffer(("send", j), sz[j], device)
            pack_fn(key, buf)
            ops.append(dist.P2POp(dist.isend, buf, dst))
            moved += sz[j]
        elif rank == dst:
            buf = _buffer(("recv", j), sz[j], device)
            ops.append(dist.P2POp(dist.irecv, buf, src))
            incoming.append((key, buf, sz[j]))
            moved += sz[j]
    if ops:
        if torch.device(device).type == "cuda":
            torch.cuda.synchronize(device)      # the engine packed on its own stream
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        if torch.device(device).type == "cuda":
            torch.cuda.synchronize(device)      # NCCL is asynchronous; the engine reads on its own stream
    for key, buf, nb in incoming:
        unpack_fn(key, buf, nb)
    return moved


def exchange_to_root(items, rank, nbytes_fn, pack_fn, unpack_fn, device, root=0):
    """The single-level special case: items (key, owner) all travel to `root`."""
    return exchange([(k, o, root) for k, o in items], rank, nbytes_fn, pack_fn, unpack_fn, device)


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
        self.debug = False

    def analyse(self, n, ptr, row, order):
        flag = self.s.analyse(n, ptr, row, order)
        self.n = n
        s = self.s
        self.rank_struct = int(s.inform.matrix_rank)
        f = s.L.b200ldlt_dist_init(s.akeep, C.byref(s.fkeep), self.rank, self.world, C.byref(s.opt), C.byref(s.inform))
        assert f >= 0, f"dist_init flag {f}"
        self.nstages = s.L.b200ldlt_dist_nstages(s.akeep)
        self.boundaries = [[] for _ in range(self.nstages)]      # per stage: (node, src rank, dst rank)
        for i in range(s.L.b200ldlt_dist_nboundary(s.akeep)):
            node, src, dst, stage = C.c_int(), C.c_int(), C.c_int(), C.c_int()
            s.L.b200ldlt_dist_boundary(s.akeep, i, C.byref(node), C.byref(src), C.byref(dst), C.byref(stage))
            self.boundaries[stage.value].append((node.value, src.value, dst.value))
        return flag

    def _agree(self, flag):
        """The smallest flag over the ranks (errors are negative): every rank takes the same decision."""
        if self.world == 1:
            return int(flag)
        t = torch.tensor([int(flag)], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return int(t.item())

    def factor(self, val, posdef=False):
        """Returns the SSIDS flag of the factorisation: the most severe error of any rank (MIN over the ranks),
        else the largest warning (MAX).  An error in a stage stops every rank before the exchange of that stage."""
        s = self.s
        val = np.ascontiguousarray(val, np.float64)
        vp = val.ctypes.data_as(C.POINTER(C.c_double))
        self.bytes_moved = 0
        self.t_stage = []
        self.launches = 0
        warn = 0
        for k in range(self.nstages):
            f = s.L.b200ldlt_dist_factor(s.akeep, s.fkeep, k + 1, int(posdef), vp, C.byref(s.opt), C.byref(s.inform))
            self.t_stage.append(s.inform.time_factor)
            self.launches += int(s.inform.reserved[0])
            worst = self._agree(min(f, 0))
            if worst < 0:
                self.num_neg = self.num_two = self.num_delay = self.num_zero = 0
                return worst
            warn = max(warn, f)
            self.bytes_moved += exchange(
                self.boundaries[k], self.rank,
                lambda node: s.L.b200ldlt_dist_contrib_bytes(s.fkeep, node),
                lambda node, buf: s.L.b200ldlt_dist_contrib_pack(s.fkeep, node, C.c_void_p(buf.data_ptr())),
                lambda node, buf, nb: s.L.b200ldlt_dist_contrib_unpack(s.fkeep, node, C.c_void_p(buf.data_ptr()), nb),
                self.device)
        self.t_leaves, self.t_root = self.t_stage[0], float(sum(self.t_stage[1:]))
        # after its last phase a rank's counters cover every front it factorised; zero pivots are the difference between
        # the structural rank of the analysis and the rank the engine reports
        nzero = int(self.rank_struct - s.inform.matrix_rank)
        stats = np.array([s.inform.num_neg, s.inform.num_two, s.inform.num_delay, nzero], dtype=np.int64)
        t = torch.from_numpy(stats).to(self.device)
        w = torch.tensor([int(warn)], dtype=torch.int64, device=self.device)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            dist.all_reduce(w, op=dist.ReduceOp.MAX)
        tot = t.cpu().numpy()
        self.num_neg, self.num_two, self.num_delay, self.num_zero = int(tot[0]), int(tot[1]), int(tot[2]), int(tot[3])
        self.matrix_rank = int(self.rank_struct - self.num_zero)
        warn = int(w.item())
        if self.num_zero > 0:
            warn = max(warn, 7)          # SSIDS_WARNING_FACT_SINGULAR (datatypes.f90:56), whoever found the zero pivots
        return warn

    def solve(self, b):
        s = self.s
        n = self.n
        x = torch.from_numpy(np.ascontiguousarray(b, np.float64)).to(self.device)
        torch.cuda.synchronize(self.device)
        f = s.L.b200ldlt_dist_solve(s.fkeep, 1, C.c_void_p(x.data_ptr()), None)
        assert f >= 0, f
        for k in range(self.nstages):
            f = self._agree(s.L.b200ldlt_dist_solve(s.fkeep, 100 + k, C.c_void_p(x.data_ptr()), None))
            if f < 0:
                raise RuntimeError(f"distributed forward solve failed in stage {k} on some rank: flag {f}")
            exchange(
                self.boundaries[k], self.rank,
                lambda node: s.L.b200ldlt_dist_xtail_bytes(s.fkeep, node),
                lambda node, buf: s.L.b200ldlt_dist_xtail_pack(s.fkeep, node, C.c_void_p(buf.data_ptr())),
                lambda node, buf, nb: s.L.b200ldlt_dist_xtail_unpack(s.fkeep, node, C.c_void_p(buf.data_ptr())),
                self.device)
        part = torch.empty(2 * n, dtype=torch.float64, device=self.device)
        for k in range(self.nstages - 1, -1, -1):
            f = self._agree(s.L.b200ldlt_dist_solve(s.fkeep, 200 + k, C.c_void_p(x.data_ptr()), None))
            if f < 0:
                raise RuntimeError(f"distributed backward solve failed in stage {k} on some rank: flag {f}")
            if self.world > 1:
                # the solution entries of stage k, from whoever produced them, to everybody
                assert s.L.b200ldlt_dist_xp_get_stage(s.fkeep, k, C.c_void_p(part.data_ptr()), n) == 0
                dist.all_reduce(part, op=dist.ReduceOp.SUM)
                torch.cuda.synchronize(self.device)
                assert s.L.b200ldlt_dist_xp_merge_stage(s.fkeep, k, C.c_void_p(part.data_ptr()), n) == 0
                if self.debug:
                    xpv = torch.empty(n, dtype=torch.float64, device=self.device)
                    s.L.b200ldlt_dist_xp_get(s.fkeep, C.c_void_p(xpv.data_ptr()), n)
                    print(f"[rank {self.rank}] bwd stage {k}: flagged {int(part[n:].sum().item())} sum|val| {part[:n].abs().sum().item():.6e} "
                          f"sum|xp| {xpv.abs().sum().item():.6e}", flush=True)
        f = s.L.b200ldlt_dist_solve(s.fkeep, 4, C.c_void_p(x.data_ptr()), None)
        assert f >= 0, f
        return x.cpu().numpy()
