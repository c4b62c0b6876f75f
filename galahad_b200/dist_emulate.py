"""Emulation of the multi-rank tree-parallel path inside ONE process on ONE GPU: every "rank" is a
separate fkeep on the same device and the boundary exchange is a direct device-to-device copy.
Same C-ABI calls and the same schedule as galahad_b200/dist.py; used by the -m gpu tests (the
multi-rank logic is exercised without needing several GPUs) and for debugging."""
import ctypes as C

import numpy as np
import torch

from ._lib import Solver


class EmulatedRanks:
    def __init__(self, nranks, **opts):
        self.nr = nranks
        self.s = [Solver(device=0, **opts) for _ in range(nranks)]

    def analyse(self, n, ptr, row, order):
        self.n = n
        for r, s in enumerate(self.s):
            assert s.analyse(n, ptr, row, order) >= 0
            f = s.L.b200ldlt_dist_init(s.akeep, C.byref(s.fkeep), r, self.nr, C.byref(s.opt), C.byref(s.inform))
            assert f >= 0, f
        s = self.s[0]
        self.boundaries = []
        for i in range(s.L.b200ldlt_dist_nboundary(s.akeep)):
            node, owner = C.c_int(), C.c_int()
            s.L.b200ldlt_dist_boundary(s.akeep, i, C.byref(node), C.byref(owner))
            self.boundaries.append((node.value, owner.value))

    def factor(self, val, posdef=False):
        val = np.ascontiguousarray(val, np.float64)
        vp = val.ctypes.data_as(C.POINTER(C.c_double))
        neg = two = delay = 0
        for r, s in enumerate(self.s):
            f = s.L.b200ldlt_dist_factor(s.akeep, s.fkeep, 1, int(posdef), vp, C.byref(s.opt), C.byref(s.inform))
            assert f >= 0, (r, f, s.inform.cuda_error)
            if r != 0:
                neg += s.inform.num_neg; two += s.inform.num_two; delay += s.inform.num_delay
        s0 = self.s[0]
        self.moved = 0
        for node, owner in self.boundaries:
            if owner == 0:
                continue
            so = self.s[owner]
            nb = so.L.b200ldlt_dist_contrib_bytes(so.fkeep, node)
            buf = torch.empty(max(nb, 1), dtype=torch.uint8, device="cuda")
            assert so.L.b200ldlt_dist_contrib_pack(so.fkeep, node, C.c_void_p(buf.data_ptr())) == 0
            assert s0.L.b200ldlt_dist_contrib_unpack(s0.fkeep, node, C.c_void_p(buf.data_ptr()), nb) == 0
            self.moved += nb
        f = s0.L.b200ldlt_dist_factor(s0.akeep, s0.fkeep, 2, int(posdef), vp, C.byref(s0.opt), C.byref(s0.inform))
        assert f >= 0, (f, s0.inform.cuda_error)
        self.num_neg = neg + s0.inform.num_neg
        self.num_two = two + s0.inform.num_two
        self.num_delay = delay + s0.inform.num_delay
        return f

    def solve(self, b):
        n = self.n
        xs = [torch.from_numpy(np.ascontiguousarray(b, np.float64)).cuda() for _ in self.s]
        for s, x in zip(self.s, xs):
            assert s.L.b200ldlt_dist_solve(s.fkeep, 1, C.c_void_p(x.data_ptr()), None) == 0
        s0 = self.s[0]
        for node, owner in self.boundaries:
            if owner == 0:
                continue
            so = self.s[owner]
            nb = so.L.b200ldlt_dist_xtail_bytes(so.fkeep, node)
            buf = torch.empty(max(nb, 1), dtype=torch.uint8, device="cuda")
            assert so.L.b200ldlt_dist_xtail_pack(so.fkeep, node, C.c_void_p(buf.data_ptr())) == 0
            assert s0.L.b200ldlt_dist_xtail_unpack(s0.fkeep, node, C.c_void_p(buf.data_ptr())) == 0
        assert s0.L.b200ldlt_dist_solve(s0.fkeep, 2, C.c_void_p(xs[0].data_ptr()), None) == 0
        xp = torch.zeros(n, dtype=torch.float64, device="cuda")
        s0.L.b200ldlt_dist_xp_get(s0.fkeep, C.c_void_p(xp.data_ptr()), n)
        for r, (s, x) in enumerate(zip(self.s, xs)):
            src = None if r == 0 else C.c_void_p(xp.data_ptr())
            assert s.L.b200ldlt_dist_solve(s.fkeep, 3, C.c_void_p(x.data_ptr()), src) == 0
            assert s.L.b200ldlt_dist_solve(s.fkeep, 4, C.c_void_p(x.data_ptr()), None) == 0
        tot = xs[0]
        for x in xs[1:]:
            tot = tot + x
        return tot.cpu().numpy()

    def free(self):
        for s in self.s:
            s.free()
