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
        self.nstages = s.L.b200ldlt_dist_nstages(s.akeep)
        self.boundaries = [[] for _ in range(self.nstages)]
        for i in range(s.L.b200ldlt_dist_nboundary(s.akeep)):
            node, src, dst, stage = C.c_int(), C.c_int(), C.c_int(), C.c_int()
            s.L.b200ldlt_dist_boundary(s.akeep, i, C.byref(node), C.byref(src), C.byref(dst), C.byref(stage))
            self.boundaries[stage.value].append((node.value, src.value, dst.value))

    def factor(self, val, posdef=False):
        val = np.ascontiguousarray(val, np.float64)
        vp = val.ctypes.data_as(C.POINTER(C.c_double))
        self.moved = 0
        for k in range(self.nstages):
            for r, s in enumerate(self.s):
                f = s.L.b200ldlt_dist_factor(s.akeep, s.fkeep, k + 1, int(posdef), vp, C.byref(s.opt), C.byref(s.inform))
                assert f >= 0, (k, r, f, s.inform.cuda_error)
            for node, src, dst in self.boundaries[k]:
                if src == dst:
                    continue
                so, sd = self.s[src], self.s[dst]
                nb = so.L.b200ldlt_dist_contrib_bytes(so.fkeep, node)
                buf = torch.empty(max(nb, 1), dtype=torch.uint8, device="cuda")
                assert so.L.b200ldlt_dist_contrib_pack(so.fkeep, node, C.c_void_p(buf.data_ptr())) == 0
                assert sd.L.b200ldlt_dist_contrib_unpack(sd.fkeep, node, C.c_void_p(buf.data_ptr()), nb) == 0
                self.moved += nb
        # after its last phase a rank's counters cover every front it factorised
        self.num_neg = sum(s.inform.num_neg for s in self.s)
        self.num_two = sum(s.inform.num_two for s in self.s)
        self.num_delay = sum(s.inform.num_delay for s in self.s)
        return 0

    def solve(self, b):
        n = self.n
        xs = [torch.from_numpy(np.ascontiguousarray(b, np.float64)).cuda() for _ in self.s]
        for s, x in zip(self.s, xs):
            assert s.L.b200ldlt_dist_solve(s.fkeep, 1, C.c_void_p(x.data_ptr()), None) == 0
        for k in range(self.nstages):
            for s, x in zip(self.s, xs):
                assert s.L.b200ldlt_dist_solve(s.fkeep, 100 + k, C.c_void_p(x.data_ptr()), None) == 0
            for node, src, dst in self.boundaries[k]:
                if src == dst:
                    continue
                so, sd = self.s[src], self.s[dst]
                nb = so.L.b200ldlt_dist_xtail_bytes(so.fkeep, node)
                buf = torch.empty(max(nb, 1), dtype=torch.uint8, device="cuda")
                assert so.L.b200ldlt_dist_xtail_pack(so.fkeep, node, C.c_void_p(buf.data_ptr())) == 0
                assert sd.L.b200ldlt_dist_xtail_unpack(sd.fkeep, node, C.c_void_p(buf.data_ptr())) == 0
        for k in range(self.nstages - 1, -1, -1):
            tot = torch.zeros(2 * n, dtype=torch.float64, device="cuda")
            for s, x in zip(self.s, xs):
                assert s.L.b200ldlt_dist_solve(s.fkeep, 200 + k, C.c_void_p(x.data_ptr()), None) == 0
                part = torch.empty(2 * n, dtype=torch.float64, device="cuda")
                assert s.L.b200ldlt_dist_xp_get_stage(s.fkeep, k, C.c_void_p(part.data_ptr()), n) == 0
                tot += part                                            # the all-reduce of dist.py
            torch.cuda.synchronize()
            for s in self.s:
                assert s.L.b200ldlt_dist_xp_merge_stage(s.fkeep, k, C.c_void_p(tot.data_ptr()), n) == 0
        out = []
        for s, x in zip(self.s, xs):
            assert s.L.b200ldlt_dist_solve(s.fkeep, 4, C.c_void_p(x.data_ptr()), None) == 0
            out.append(x.cpu().numpy())
        for o in out[1:]:
            assert np.array_equal(o, out[0])                          # every rank ends with the whole solution
        return out[0]

    def free(self):
        for s in self.s:
            s.free()
