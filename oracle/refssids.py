"""TEST INFRASTRUCTURE ONLY (oracle).  ctypes wrappers around

  * oracle/_ref/libssids_ref.so  -- the reference's own C++ SSIDS CPU engine, compiled unmodified
    from /root/reference by oracle/build_ref.sh (kind "reference"), and
  * oracle/liboracle.so          -- our plain-C restatement oracle/ldlt_oracle.c (kind "port").

Both take the symbolic arrays (sptr, sparent, rptr, rlist, nptr, nlist) exactly as analyse produced
them, plus A's values, and work on right-hand sides held in PIVOT ORDER.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libssids_ref.so")
ORACLE_SO = os.path.join(HERE, "liboracle.so")

# must be in the environment before libgomp initialises (src/ssids/ssids.f90:1439-1443)
os.environ.setdefault("OMP_CANCELLATION", "true")
os.environ.setdefault("OMP_PROC_BIND", "true")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
os.environ.setdefault("OMP_MAX_ACTIVE_LEVELS", "4")


def build_oracle(force=False):
    src = os.path.join(HERE, "ldlt_oracle.c")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["/usr/bin/gcc", "-O2", "-fPIC", "-shared", "-o", ORACLE_SO, src, "-lm"])
    return ORACLE_SO


def build_ref():
    subprocess.check_call(["bash", os.path.join(HERE, "build_ref.sh")])
    return REF_SO


def have_ref():
    return os.path.exists(REF_SO)


class _Opts(C.Structure):
    _fields_ = [("print_level", C.c_int), ("action", C.c_int), ("small_", C.c_double), ("u", C.c_double),
                ("multiplier", C.c_double), ("small_subtree_threshold", C.c_long), ("cpu_block_size", C.c_int),
                ("pivot_method", C.c_int), ("failed_pivot_method", C.c_int)]


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


STAT_NAMES = ["flag", "num_delay", "num_neg", "num_two", "num_zero", "maxfront", "not_first_pass", "not_second_pass"]


class RefSSIDS:
    """The reference engine (single part, all threads), options = ssids_options defaults."""

    def __init__(self, sym, n, u=0.01, small=1e-20, action=True, pivot_method=2):
        self.lib = C.CDLL(REF_SO)
        L = self.lib
        L.refssids_symbolic.restype = C.c_void_p
        L.refssids_factor.restype = C.c_void_p
        L.refssids_symbolic.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_long),
                                        C.POINTER(C.c_int), C.POINTER(C.c_long), C.POINTER(C.c_long), C.POINTER(_Opts)]
        L.refssids_factor.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_void_p, C.POINTER(_Opts),
                                      C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.refssids_solve.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int]
        L.refssids_enquire.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.refssids_alter.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.refssids_free_numeric.argtypes = [C.c_void_p, C.c_int]
        L.refssids_free_symbolic.argtypes = [C.c_void_p]
        self.n = n
        self.opts = _Opts(0, int(action), small, u, 1.1, 4000000, 256, pivot_method, 1)
        # 1-based copies, kept alive (the engine keeps pointers into them)
        self.sptr = (sym["sptr"] + 1).astype(np.int32)
        self.sparent = (sym["sparent"] + 1).astype(np.int32)
        self.rptr = (sym["rptr"] + 1).astype(np.int64)
        self.rlist = (sym["rlist"] + 1).astype(np.int32)
        self.nptr = (sym["nptr"] + 1).astype(np.int64)
        self.nlist = (sym["nlist"] + 1).astype(np.int64)
        self.nnodes = int(sym["nnodes"])
        self.symb = L.refssids_symbolic(n, self.nnodes, _p(self.sptr, C.c_int), _p(self.sparent, C.c_int),
                                        _p(self.rptr, C.c_long), _p(self.rlist, C.c_int), _p(self.nptr, C.c_long),
                                        _p(self.nlist, C.c_long), C.byref(self.opts))
        self.num = None
        self.posdef = False

    def factor(self, val, posdef=False, nthreads=0):
        if self.num:
            self.lib.refssids_free_numeric(self.num, int(self.posdef))
            self.num = None
        self.val = np.ascontiguousarray(val, np.float64)
        stats = np.zeros(8, np.int32)
        secs = C.c_double()
        self.posdef = bool(posdef)
        self.num = self.lib.refssids_factor(self.symb, int(posdef), _p(self.val, C.c_double), None, C.byref(self.opts),
                                            nthreads, _p(stats, C.c_int), C.byref(secs))
        st = dict(zip(STAT_NAMES, [int(v) for v in stats]))
        st["seconds"] = secs.value
        return st

    def solve(self, x_piv, job=0):
        x = np.array(x_piv, np.float64, order="F", copy=True)
        nrhs = 1 if x.ndim == 1 else x.shape[1]
        f = self.lib.refssids_solve(self.num, int(self.posdef), job, nrhs, _p(x, C.c_double), x.shape[0])
        assert f >= 0, f
        return x

    def enquire(self):
        piv = np.zeros(self.n, np.int32)
        d = np.zeros(2 * self.n)
        self.lib.refssids_enquire(self.num, int(self.posdef), _p(piv, C.c_int), _p(d, C.c_double))
        return piv, d

    def nthreads(self):
        return self.lib.refssids_max_threads()

    def close(self):
        if self.num:
            self.lib.refssids_free_numeric(self.num, int(self.posdef))
            self.num = None
        if self.symb:
            self.lib.refssids_free_symbolic(self.symb)
            self.symb = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Oracle:
    """Our C restatement (oracle/ldlt_oracle.c); 0-based arrays."""

    def __init__(self, sym, n, u=0.01, small=1e-20, action=True):
        build_oracle()
        self.lib = C.CDLL(ORACLE_SO)
        L = self.lib
        L.oracle_factor.restype = C.c_void_p
        L.oracle_factor.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64),
                                    C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_double),
                                    C.c_int, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_int)]
        L.oracle_solve.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int]
        L.oracle_enquire.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_node_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_node_L.restype = C.POINTER(C.c_double)
        L.oracle_node_L.argtypes = [C.c_void_p, C.c_int]
        L.oracle_node_d.restype = C.POINTER(C.c_double)
        L.oracle_node_d.argtypes = [C.c_void_p, C.c_int]
        L.oracle_node_perm.restype = C.POINTER(C.c_int)
        L.oracle_node_perm.argtypes = [C.c_void_p, C.c_int]
        self.n, self.u, self.small, self.action = n, u, small, int(action)
        self.sym = {k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in sym.items()}
        self.f = None

    def factor(self, val, posdef=False):
        if self.f:
            self.lib.oracle_free(self.f)
        s = self.sym
        self.val = np.ascontiguousarray(val, np.float64)
        stats = np.zeros(8, np.int32)
        self.f = self.lib.oracle_factor(self.n, int(s["nnodes"]), _p(s["sptr"], C.c_int), _p(s["sparent"], C.c_int),
                                        _p(s["rptr"], C.c_int64), _p(s["rlist"], C.c_int), _p(s["nptr"], C.c_int64),
                                        _p(s["nlist"], C.c_int64), _p(self.val, C.c_double), int(posdef), self.u,
                                        self.small, self.action, _p(stats, C.c_int))
        return dict(zip(STAT_NAMES, [int(v) for v in stats]))

    def solve(self, x_piv, job=0):
        x = np.array(x_piv, np.float64, order="F", copy=True)
        nrhs = 1 if x.ndim == 1 else x.shape[1]
        f = self.lib.oracle_solve(self.f, job, nrhs, _p(x, C.c_double), x.shape[0])
        assert f >= 0, f
        return x

    def enquire(self):
        piv = np.zeros(self.n, np.int32)
        d = np.zeros(2 * self.n)
        self.lib.oracle_enquire(self.f, _p(piv, C.c_int), _p(d, C.c_double))
        return piv, d

    def node(self, i):
        m, n, ne = C.c_int(), C.c_int(), C.c_int()
        self.lib.oracle_node_info(self.f, i, C.byref(m), C.byref(n), C.byref(ne))
        m, n, ne = m.value, n.value, ne.value
        Lp = self.lib.oracle_node_L(self.f, i)
        L = np.ctypeslib.as_array(Lp, shape=(max(ne, 1), max(m, 1)))[:ne, :m].T.copy() if ne > 0 else np.zeros((m, 0))
        d = np.ctypeslib.as_array(self.lib.oracle_node_d(self.f, i), shape=(2 * max(n, 1),))[:2 * n].copy()
        perm = np.ctypeslib.as_array(self.lib.oracle_node_perm(self.f, i), shape=(max(n, 1),))[:n].copy()
        return dict(m=m, n=n, nelim=ne, L=L, d=d, perm=perm)

    def close(self):
        if self.f:
            self.lib.oracle_free(self.f)
            self.f = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def permuted_solve(engine, order, b, job=0):
    """Apply the pivot-order permutation around an oracle solve: x(order) as fkeep.F90:248-310."""
    b = np.asarray(b, np.float64)
    xp = np.zeros_like(b)
    xp[order] = b
    xp = engine.solve(xp, job)
    return xp[order]
