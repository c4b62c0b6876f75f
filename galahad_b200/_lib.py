"""ctypes binding of the C ABI in include/b200ldlt.h (libb200ldlt.so, built in-tree by build.py).

There is no Python or CPU implementation behind these calls: if the shared library is missing the
import fails loudly, and on a box without a CUDA device factor/solve return flag -51.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200ldlt.so")


class Options(C.Structure):
    _fields_ = [
        ("array_base", C.c_int),
        ("print_level", C.c_int),
        ("nemin", C.c_int),
        ("action", C.c_int),
        ("u", C.c_double),
        ("small_", C.c_double),
        ("multiplier", C.c_double),
        ("device", C.c_int),
        ("panel_width", C.c_int),
        ("reserved", C.c_int * 7),
    ]


class Inform(C.Structure):
    _fields_ = [
        ("flag", C.c_int),
        ("matrix_dup", C.c_int),
        ("matrix_missing_diag", C.c_int),
        ("matrix_outrange", C.c_int),
        ("matrix_rank", C.c_int),
        ("maxdepth", C.c_int),
        ("maxfront", C.c_int),
        ("num_delay", C.c_int),
        ("num_factor", C.c_int64),
        ("num_flops", C.c_int64),
        ("num_neg", C.c_int),
        ("num_sup", C.c_int),
        ("num_two", C.c_int),
        ("stat", C.c_int),
        ("cuda_error", C.c_int),
        ("nparts", C.c_int),
        ("not_first_pass", C.c_int),
        ("not_second_pass", C.c_int),
        ("time_analyse", C.c_double),
        ("time_factor", C.c_double),
        ("time_factor_device", C.c_double),
        ("time_solve", C.c_double),
        ("factor_bytes", C.c_int64),
        ("upd_flops", C.c_double),
        ("upd_seconds", C.c_double),
        ("reserved", C.c_int * 8),
    ]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


# every symbol include/b200ldlt.h declares (tests check the library exports each of them)
SYMBOLS = [
    "b200ldlt_default_options", "b200ldlt_order", "b200ldlt_analyse", "b200ldlt_analyse_ptr64",
    "b200ldlt_analyse_matched", "b200ldlt_match_order", "b200ldlt_factor",
    "b200ldlt_factor_device", "b200ldlt_set_value_map", "b200ldlt_factor_mapped", "b200ldlt_factor_changed", "b200ldlt_set_user_matrix", "b200ldlt_solve_refined", "b200ldlt_factor_scaled", "b200ldlt_equilib_scale", "b200ldlt_hungarian_scale", "b200ldlt_auction_scale", "b200ldlt_set_scaling", "b200ldlt_solve", "b200ldlt_solve_device", "b200ldlt_enquire_indef",
    "b200ldlt_enquire_posdef", "b200ldlt_alter", "b200ldlt_free_fkeep", "b200ldlt_free",
    "b200ldlt_akeep_sizes", "b200ldlt_akeep_get", "b200ldlt_partition", "b200ldlt_part_info", "b200ldlt_map_info",
    "b200ldlt_node_info", "b200ldlt_node_copy", "b200ldlt_version", "b200ldlt_ntree_sizes", "b200ldlt_ntree_get",
    "b200ldlt_dist_init", "b200ldlt_dist_factor", "b200ldlt_dist_nstages", "b200ldlt_dist_nboundary", "b200ldlt_dist_boundary",
    "b200ldlt_dist_xp_get_stage", "b200ldlt_dist_xp_merge_stage",
    "b200ldlt_dist_contrib_bytes", "b200ldlt_dist_contrib_pack", "b200ldlt_dist_contrib_unpack",
    "b200ldlt_dist_xtail_bytes", "b200ldlt_dist_xtail_pack", "b200ldlt_dist_xtail_unpack", "b200ldlt_dist_solve",
    "b200ldlt_dist_xp", "b200ldlt_dist_xp_get",
]

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m galahad_b200.build` "
            "(nvcc, sm_100a). b200ldlt has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, vpp = C.c_void_p, C.POINTER(C.c_void_p)
    ip, dp, lp = C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_int64)
    op, fp = C.POINTER(Options), C.POINTER(Inform)
    L.b200ldlt_default_options.argtypes = [op]
    L.b200ldlt_default_options.restype = None
    L.b200ldlt_order.argtypes = [C.c_int, lp, ip, C.c_int, ip, op]
    L.b200ldlt_analyse.argtypes = [vpp, C.c_int, ip, ip, ip, op, fp]
    L.b200ldlt_analyse_ptr64.argtypes = [vpp, C.c_int, lp, ip, ip, op, fp]
    L.b200ldlt_analyse_matched.argtypes = [vpp, C.c_int, lp, ip, dp, ip, op, fp]
    L.b200ldlt_match_order.argtypes = [C.c_int, lp, ip, dp, C.c_int, ip, dp, ip, ip, lp, ip]
    L.b200ldlt_factor.argtypes = [vp, vpp, C.c_int, dp, op, fp]
    L.b200ldlt_factor_device.argtypes = [vp, vpp, C.c_int, vp, op, fp]
    L.b200ldlt_set_value_map.argtypes = [vp, vpp, C.c_int, ip, op, fp]
    L.b200ldlt_factor_mapped.argtypes = [vp, vpp, C.c_int, C.c_int, dp, op, fp]
    L.b200ldlt_factor_changed.argtypes = [vp, vpp, C.c_int, C.c_int, ip, dp, op, fp]
    L.b200ldlt_set_user_matrix.argtypes = [vp, vpp, C.c_int, ip, ip, op, fp]
    L.b200ldlt_factor_scaled.argtypes = [vp, vpp, C.c_int, dp, dp, op, fp]
    L.b200ldlt_equilib_scale.argtypes = [C.c_int, lp, ip, dp, dp, ip]
    L.b200ldlt_hungarian_scale.argtypes = [C.c_int, lp, ip, dp, dp, ip, ip, C.c_int]
    L.b200ldlt_auction_scale.argtypes = [C.c_int, lp, ip, dp, dp, ip, ip, ip]
    L.b200ldlt_set_scaling.argtypes = [vp, vpp, dp, op, fp]
    L.b200ldlt_solve_refined.argtypes = [vp, vp, dp, C.c_int, C.c_double, C.c_double, ip, op, fp]
    L.b200ldlt_solve.argtypes = [vp, vp, C.c_int, C.c_int, dp, C.c_int, op, fp]
    L.b200ldlt_solve_device.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, op, fp]
    L.b200ldlt_enquire_indef.argtypes = [vp, vp, ip, dp, op, fp]
    L.b200ldlt_enquire_posdef.argtypes = [vp, vp, dp, fp]
    L.b200ldlt_alter.argtypes = [vp, vp, dp, fp]
    L.b200ldlt_free_fkeep.argtypes = [vpp]
    L.b200ldlt_free.argtypes = [vpp, vpp]
    L.b200ldlt_akeep_sizes.argtypes = [vp, ip, lp, lp]
    L.b200ldlt_akeep_get.argtypes = [vp, C.c_int, ip, ip, lp, ip, lp, lp, ip]
    L.b200ldlt_ntree_sizes.argtypes = [vp, ip, lp, lp]
    L.b200ldlt_ntree_get.argtypes = [vp, C.c_int, ip, ip, lp, ip, lp, lp]
    L.b200ldlt_dist_init.argtypes = [vp, vpp, C.c_int, C.c_int, op, fp]
    L.b200ldlt_dist_factor.argtypes = [vp, vp, C.c_int, C.c_int, dp, op, fp]
    L.b200ldlt_dist_nboundary.argtypes = [vp]
    L.b200ldlt_dist_boundary.argtypes = [vp, C.c_int, ip, ip, ip, ip]
    L.b200ldlt_dist_nstages.argtypes = [vp]
    L.b200ldlt_dist_xp_get_stage.argtypes = [vp, C.c_int, vp, C.c_int]
    L.b200ldlt_dist_xp_merge_stage.argtypes = [vp, C.c_int, vp, C.c_int]
    L.b200ldlt_dist_contrib_bytes.argtypes = [vp, C.c_int]
    L.b200ldlt_dist_contrib_bytes.restype = C.c_int64
    L.b200ldlt_dist_contrib_pack.argtypes = [vp, C.c_int, vp]
    L.b200ldlt_dist_contrib_unpack.argtypes = [vp, C.c_int, vp, C.c_int64]
    L.b200ldlt_dist_xtail_bytes.argtypes = [vp, C.c_int]
    L.b200ldlt_dist_xtail_bytes.restype = C.c_int64
    L.b200ldlt_dist_xtail_pack.argtypes = [vp, C.c_int, vp]
    L.b200ldlt_dist_xtail_unpack.argtypes = [vp, C.c_int, vp]
    L.b200ldlt_dist_solve.argtypes = [vp, C.c_int, vp, vp]
    L.b200ldlt_dist_xp.argtypes = [vp]
    L.b200ldlt_dist_xp.restype = C.c_void_p
    L.b200ldlt_dist_xp_get.argtypes = [vp, vp, C.c_int]
    L.b200ldlt_partition.argtypes = [vp, C.c_int]
    L.b200ldlt_map_info.argtypes = [vp, C.c_int, ip, ip]
    L.b200ldlt_part_info.argtypes = [vp, C.c_int, ip, ip, ip, ip, lp]
    L.b200ldlt_node_info.argtypes = [vp, C.c_int, ip, ip, ip]
    L.b200ldlt_node_copy.argtypes = [vp, C.c_int, dp, dp, ip]
    L.b200ldlt_version.restype = C.c_char_p
    for s in SYMBOLS:
        getattr(L, s)
    _lib = L
    return L


def as_ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def default_options(**kw):
    o = Options()
    lib().b200ldlt_default_options(C.byref(o))
    for k, v in kw.items():
        if k == "small":
            k = "small_"
        setattr(o, k, v)
    return o


class Solver:
    """Handle pair (akeep, fkeep) with numpy-friendly methods; 0-based indices."""

    def __init__(self, **opts):
        self.L = lib()
        self.opt = default_options(**opts)
        self.akeep = C.c_void_p()
        self.fkeep = C.c_void_p()
        self.inform = Inform()
        self.n = 0

    def analyse(self, n, ptr, row, order):
        ptr = np.ascontiguousarray(ptr, dtype=np.int64)
        row = np.ascontiguousarray(row, dtype=np.int32)
        self.order = np.ascontiguousarray(order, dtype=np.int32).copy()
        self.n = int(n)
        flag = self.L.b200ldlt_analyse_ptr64(C.byref(self.akeep), self.n, as_ptr(ptr, C.c_int64),
                                             as_ptr(row, C.c_int), as_ptr(self.order, C.c_int),
                                             C.byref(self.opt), C.byref(self.inform))
        return flag

    def analyse_matched(self, n, ptr, row, val):
        """SSIDS options%ordering = 2: matching-based order from the values; its scaling is kept for factor_scaled(scaling=3)."""
        ptr = np.ascontiguousarray(ptr, dtype=np.int64)
        row = np.ascontiguousarray(row, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        self.order = np.zeros(max(int(n), 1), np.int32)[:int(n)]
        self.n = int(n)
        return self.L.b200ldlt_analyse_matched(C.byref(self.akeep), self.n, as_ptr(ptr, C.c_int64), as_ptr(row, C.c_int),
                                               as_ptr(val, C.c_double), as_ptr(self.order, C.c_int),
                                               C.byref(self.opt), C.byref(self.inform))

    def symbolic(self, base=0):
        nn, nr, nl = C.c_int(), C.c_int64(), C.c_int64()
        self.L.b200ldlt_akeep_sizes(self.akeep, C.byref(nn), C.byref(nr), C.byref(nl))
        nn, nr, nl = nn.value, nr.value, nl.value
        out = dict(
            nnodes=nn,
            sptr=np.zeros(nn + 1, np.int32), sparent=np.zeros(max(nn, 1), np.int32)[:nn],
            rptr=np.zeros(nn + 1, np.int64), rlist=np.zeros(max(nr, 1), np.int32)[:nr],
            nptr=np.zeros(nn + 1, np.int64), nlist=np.zeros(max(2 * nl, 1), np.int64)[:2 * nl],
            invp=np.zeros(max(self.n, 1), np.int32)[:self.n])
        self.L.b200ldlt_akeep_get(self.akeep, base, as_ptr(out["sptr"], C.c_int), as_ptr(out["sparent"], C.c_int),
                                  as_ptr(out["rptr"], C.c_int64), as_ptr(out["rlist"], C.c_int),
                                  as_ptr(out["nptr"], C.c_int64), as_ptr(out["nlist"], C.c_int64),
                                  as_ptr(out["invp"], C.c_int))
        return out

    def numeric_tree(self, base=0):
        """The engine's own tree (chains fused); same keys as symbolic() minus invp."""
        nn, nr, nl = C.c_int(), C.c_int64(), C.c_int64()
        self.L.b200ldlt_ntree_sizes(self.akeep, C.byref(nn), C.byref(nr), C.byref(nl))
        nn, nr, nl = nn.value, nr.value, nl.value
        out = dict(
            nnodes=nn,
            sptr=np.zeros(nn + 1, np.int32), sparent=np.zeros(max(nn, 1), np.int32)[:nn],
            rptr=np.zeros(nn + 1, np.int64), rlist=np.zeros(max(nr, 1), np.int32)[:nr],
            nptr=np.zeros(nn + 1, np.int64), nlist=np.zeros(max(2 * nl, 1), np.int64)[:2 * nl])
        self.L.b200ldlt_ntree_get(self.akeep, base, as_ptr(out["sptr"], C.c_int), as_ptr(out["sparent"], C.c_int),
                                  as_ptr(out["rptr"], C.c_int64), as_ptr(out["rlist"], C.c_int),
                                  as_ptr(out["nptr"], C.c_int64), as_ptr(out["nlist"], C.c_int64))
        return out

    def factor(self, val, posdef=False):
        val = np.ascontiguousarray(val, dtype=np.float64)
        return self.L.b200ldlt_factor(self.akeep, C.byref(self.fkeep), int(posdef), as_ptr(val, C.c_double),
                                      C.byref(self.opt), C.byref(self.inform))

    def set_value_map(self, maps):
        """maps: SLS's MAPS (1-based signed positions, sls.f90:8409-8578) for the caller's entries."""
        maps = np.ascontiguousarray(maps, dtype=np.int32)
        self._map_ne = maps.size
        return self.L.b200ldlt_set_value_map(self.akeep, C.byref(self.fkeep), maps.size, as_ptr(maps, C.c_int),
                                             C.byref(self.opt), C.byref(self.inform))

    def factor_mapped(self, user_val, posdef=False):
        user_val = np.ascontiguousarray(user_val, dtype=np.float64)
        assert user_val.size == self._map_ne
        return self.L.b200ldlt_factor_mapped(self.akeep, C.byref(self.fkeep), int(posdef), user_val.size,
                                             as_ptr(user_val, C.c_double), C.byref(self.opt), C.byref(self.inform))

    def factor_changed(self, which, new_val, posdef=False):
        """Refactorise after entries `which` (positions in the caller's list given to factor_mapped, options%array_base)
        took the values `new_val`; everything else stays as the latest factor_mapped left it on the device."""
        which = np.ascontiguousarray(which, dtype=np.int32)
        new_val = np.ascontiguousarray(new_val, dtype=np.float64)
        assert which.size == new_val.size
        return self.L.b200ldlt_factor_changed(self.akeep, C.byref(self.fkeep), int(posdef), which.size, as_ptr(which, C.c_int),
                                              as_ptr(new_val, C.c_double), C.byref(self.opt), C.byref(self.inform))

    def factor_scaled(self, val, scale=None, scaling=0, posdef=False):
        """scaling: SSIDS options%scaling (0 with `scale` = user-supplied factors, 1 = Hungarian, 4 = equilibration).
        Returns (flag, scale used or None)."""
        val = np.ascontiguousarray(val, dtype=np.float64)
        self.opt.reserved[2] = scaling
        sc = None
        if scale is not None:
            sc = np.array(scale, dtype=np.float64)
        elif scaling >= 1:
            sc = np.zeros(max(self.n, 1))
        flag = self.L.b200ldlt_factor_scaled(self.akeep, C.byref(self.fkeep), int(posdef), as_ptr(val, C.c_double),
                                             None if sc is None else as_ptr(sc, C.c_double), C.byref(self.opt), C.byref(self.inform))
        return flag, sc

    def factor_device(self, d_ptr, posdef=False):
        return self.L.b200ldlt_factor_device(self.akeep, C.byref(self.fkeep), int(posdef), C.c_void_p(d_ptr),
                                             C.byref(self.opt), C.byref(self.inform))

    def solve(self, b, job=0):
        x = np.array(b, dtype=np.float64, order="F", copy=True)
        nrhs = 1 if x.ndim == 1 else x.shape[1]
        ldx = x.shape[0]
        flag = self.L.b200ldlt_solve(self.akeep, self.fkeep, job, nrhs, as_ptr(x, C.c_double), ldx,
                                     C.byref(self.opt), C.byref(self.inform))
        return flag, x

    def solve_device(self, d_ptr, nrhs=1, ldx=None, job=0):
        return self.L.b200ldlt_solve_device(self.akeep, self.fkeep, job, nrhs, C.c_void_p(d_ptr),
                                            self.n if ldx is None else ldx, C.byref(self.opt), C.byref(self.inform))

    def enquire_indef(self):
        piv = np.zeros(max(self.n, 1), np.int32)
        d = np.zeros(2 * max(self.n, 1))
        flag = self.L.b200ldlt_enquire_indef(self.akeep, self.fkeep, as_ptr(piv, C.c_int), as_ptr(d, C.c_double),
                                             C.byref(self.opt), C.byref(self.inform))
        return flag, piv[:self.n], d[:2 * self.n].reshape(-1, 2)

    def enquire_posdef(self):
        d = np.zeros(max(self.n, 1))
        flag = self.L.b200ldlt_enquire_posdef(self.akeep, self.fkeep, as_ptr(d, C.c_double), C.byref(self.inform))
        return flag, d[:self.n]

    def alter(self, d):
        d = np.ascontiguousarray(d, dtype=np.float64).reshape(-1)
        return self.L.b200ldlt_alter(self.akeep, self.fkeep, as_ptr(d, C.c_double), C.byref(self.inform))

    def node(self, i):
        m, n, ne = C.c_int(), C.c_int(), C.c_int()
        if self.L.b200ldlt_node_info(self.fkeep, i, C.byref(m), C.byref(n), C.byref(ne)) != 0:
            return None
        m, n, ne = m.value, n.value, ne.value
        Lm = np.zeros((max(ne, 1), max(m, 1)))
        d = np.zeros(2 * max(n, 1))
        perm = np.zeros(max(n, 1), np.int32)
        if self.L.b200ldlt_node_copy(self.fkeep, i, as_ptr(Lm, C.c_double), as_ptr(d, C.c_double), as_ptr(perm, C.c_int)) != 0:
            return None
        return dict(m=m, n=n, nelim=ne, L=Lm[:ne, :m].T, d=d[:2 * n], perm=perm[:n])

    def free(self):
        self.L.b200ldlt_free(C.byref(self.akeep), C.byref(self.fkeep))

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
