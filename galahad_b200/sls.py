"""ctypes view of the SLS C interface for solver 'b200ldlt' (include/b200_sls.h).

Same call sequence as GALAHAD's C interface (include/sls.h:915-1246; test src/sls/C/slst.c):
sls_initialize -> sls_analyse_matrix -> sls_factorize_matrix -> sls_solve_system /
sls_partial_solve_system -> sls_information -> sls_terminate.
"""
import ctypes as C

import numpy as np

from ._lib import Inform as SolverInform
from ._lib import lib


class SlsControl(C.Structure):
    _fields_ = [
        ("f_indexing", C.c_bool), ("error", C.c_int), ("warning", C.c_int), ("out", C.c_int), ("statistics", C.c_int),
        ("print_level", C.c_int), ("print_level_solver", C.c_int), ("node_amalgamation", C.c_int),
        ("pivot_control", C.c_int), ("ordering", C.c_int), ("scaling", C.c_int),
        ("relative_pivot_tolerance", C.c_double), ("absolute_pivot_tolerance", C.c_double), ("zero_tolerance", C.c_double),
        ("max_iterative_refinements", C.c_int), ("acceptable_residual_relative", C.c_double),
        ("acceptable_residual_absolute", C.c_double), ("multiple_rhs", C.c_bool), ("device", C.c_int),
        ("prefix", C.c_char * 31),
    ]


class SlsTime(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("total", "analyse", "factorize", "solve", "clock_total", "clock_analyse",
                                          "clock_factorize", "clock_solve", "analyse_external", "factorize_external",
                                          "solve_external")]


class SlsInform(C.Structure):
    _fields_ = [
        ("status", C.c_int), ("alloc_status", C.c_int), ("bad_alloc", C.c_char * 81), ("more_info", C.c_int),
        ("entries", C.c_int), ("out_of_range", C.c_int), ("duplicates", C.c_int), ("upper", C.c_int),
        ("missing_diagonals", C.c_int), ("max_depth_assembly_tree", C.c_int), ("nodes_assembly_tree", C.c_int),
        ("real_size_factors", C.c_int64), ("integer_size_factors", C.c_int64), ("entries_in_factors", C.c_int64),
        ("max_front_size", C.c_int), ("two_by_two_pivots", C.c_int), ("rank", C.c_int),
        ("negative_eigenvalues", C.c_int), ("delayed_pivots", C.c_int), ("num_zero", C.c_int),
        ("iterative_refinements", C.c_int), ("flops_assembly", C.c_int64), ("flops_elimination", C.c_int64),
        ("backward_error_1", C.c_double), ("backward_error_2", C.c_double), ("time", SlsTime),
        ("b200ldlt_inform", SolverInform),
    ]


def _bind():
    L = lib()
    vpp, ip, dp = C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_double)
    cp, fp = C.POINTER(SlsControl), C.POINTER(SlsInform)
    L.sls_initialize.argtypes = [C.c_char_p, vpp, cp, ip]
    L.sls_analyse_matrix.argtypes = [cp, vpp, ip, C.c_int, C.c_char_p, C.c_int, ip, ip, ip]
    L.sls_analyse_matrix_perm.argtypes = [cp, vpp, ip, C.c_int, C.c_char_p, C.c_int, ip, ip, ip, ip]
    L.sls_reset_control.argtypes = [cp, vpp, ip]
    L.sls_factorize_matrix.argtypes = [vpp, ip, C.c_int, dp]
    L.sls_solve_system.argtypes = [vpp, ip, C.c_int, dp]
    L.sls_partial_solve_system.argtypes = [C.c_char_p, vpp, ip, C.c_int, dp]
    L.sls_information.argtypes = [vpp, fp, ip]
    L.sls_terminate.argtypes = [vpp, cp, fp]
    L.sls_coord_to_sorted_csr.argtypes = [C.c_int, C.c_int, ip, ip, ip, ip, ip, ip, ip, ip]
    for nm in ("sls_initialize", "sls_analyse_matrix", "sls_analyse_matrix_perm", "sls_reset_control",
               "sls_factorize_matrix", "sls_solve_system", "sls_partial_solve_system", "sls_information",
               "sls_terminate", "sls_coord_to_sorted_csr"):
        getattr(L, nm).restype = None
    return L


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int))


class SLS:
    """Thin object wrapper: one SLS `data` handle."""

    def __init__(self, solver="b200ldlt"):
        self.L = _bind()
        self.data = C.c_void_p()
        self.control = SlsControl()
        self.status = C.c_int()
        self.L.sls_initialize(solver.encode(), C.byref(self.data), C.byref(self.control), C.byref(self.status))

    def analyse_matrix(self, n, mtype, ne=0, row=None, col=None, ptr=None, perm=None):
        a = [None if v is None else np.ascontiguousarray(v, np.int32) for v in (row, col, ptr, perm)]
        self._keep = a
        self.n = n
        if perm is None:
            self.L.sls_analyse_matrix(C.byref(self.control), C.byref(self.data), C.byref(self.status), n,
                                      mtype.encode(), ne, _ip(a[0]), _ip(a[1]), _ip(a[2]))
        else:
            self.L.sls_analyse_matrix_perm(C.byref(self.control), C.byref(self.data), C.byref(self.status), n,
                                           mtype.encode(), ne, _ip(a[0]), _ip(a[1]), _ip(a[2]), _ip(a[3]))
        return self.status.value

    def reset_control(self):
        self.L.sls_reset_control(C.byref(self.control), C.byref(self.data), C.byref(self.status))

    def factorize_matrix(self, val):
        v = np.ascontiguousarray(val, np.float64)
        self.L.sls_factorize_matrix(C.byref(self.data), C.byref(self.status), v.size,
                                    v.ctypes.data_as(C.POINTER(C.c_double)))
        return self.status.value

    def solve_system(self, b):
        x = np.array(b, np.float64, copy=True)
        self.L.sls_solve_system(C.byref(self.data), C.byref(self.status), x.size, x.ctypes.data_as(C.POINTER(C.c_double)))
        return self.status.value, x

    def partial_solve_system(self, part, b):
        x = np.array(b, np.float64, copy=True)
        self.L.sls_partial_solve_system(part.encode(), C.byref(self.data), C.byref(self.status), x.size,
                                        x.ctypes.data_as(C.POINTER(C.c_double)))
        return self.status.value, x

    def information(self):
        inf = SlsInform()
        self.L.sls_information(C.byref(self.data), C.byref(inf), C.byref(self.status))
        return inf

    def terminate(self):
        if self.data:
            inf = SlsInform()
            self.L.sls_terminate(C.byref(self.data), C.byref(self.control), C.byref(inf))
            self.data = C.c_void_p()

    def __del__(self):
        try:
            self.terminate()
        except Exception:
            pass


def coord_to_sorted_csr(n, row, col):
    """SLS_coord_to_sorted_csr (sls.f90:8409) on 1-based coordinates; returns map, ptr, counts."""
    L = _bind()
    row = np.ascontiguousarray(row, np.int32)
    col = np.ascontiguousarray(col, np.int32)
    ne = row.size
    m = np.zeros(max(ne, 1), np.int32)
    p = np.zeros(n + 1, np.int32)
    c = [C.c_int() for _ in range(4)]
    L.sls_coord_to_sorted_csr(n, ne, _ip(row), _ip(col), _ip(m), _ip(p), *[C.byref(v) for v in c])
    return m[:ne], p, dict(dup=c[0].value, oor=c[1].value, upper=c[2].value, missing_diagonals=c[3].value)
