"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/b200ldlt.h declares, validates its arguments like ssids.f90 does, and -- with no GPU in the
box -- fails loudly (flag -51) instead of falling back to anything."""
import ctypes as C
import os
import re

import numpy as np

from conftest import gpu_available
from galahad_b200 import _lib
from galahad_b200 import problems as P

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def test_exports_every_declared_symbol():
    L = _lib.lib()
    for hdr in os.listdir(os.path.join(ROOT, "include")):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names = set(re.findall(r"\b((?:b200ldlt|b200sls|sls)_\w+)\s*\(", text))
        assert names, hdr
        for nm in names:
            assert hasattr(L, nm), f"{nm} declared in include/{hdr} but not exported"


def test_default_options_match_ssids_defaults():
    o = _lib.default_options()
    # src/ssids/datatypes.f90:21,212,254-267
    assert (o.nemin, o.action, o.u, o.small_, o.multiplier) == (32, 1, 0.01, 1e-20, 1.1)


def test_argument_checks():
    s = _lib.Solver()
    prob = P.laplace_problem(4)
    bad = prob["order"].copy()
    bad[0] = bad[1]
    assert s.analyse(prob["n"], prob["ptr"], prob["row"], bad) == -8          # SSIDS_ERROR_ORDER
    inf = _lib.Inform()
    ak = C.c_void_p()
    assert s.L.b200ldlt_analyse_ptr64(C.byref(ak), -1, None, None, None, C.byref(s.opt), C.byref(inf)) == -2
    # solve before factor: call-sequence error (ssids.f90:1166-1171)
    assert s.analyse(prob["n"], prob["ptr"], prob["row"], prob["order"]) == 0
    x = np.zeros(prob["n"])
    fk = C.c_void_p()
    assert s.L.b200ldlt_solve(s.akeep, fk, 0, 1, _lib.as_ptr(x, C.c_double), prob["n"], C.byref(s.opt), C.byref(inf)) == -1


def test_no_cpu_fallback_without_gpu():
    if gpu_available():
        return
    s = _lib.Solver()
    prob = P.laplace_problem(4)
    assert s.analyse(prob["n"], prob["ptr"], prob["row"], prob["order"]) == 0
    assert s.factor(prob["val"]) == -51                                        # SSIDS_ERROR_CUDA_UNKNOWN
    assert s.inform.cuda_error != 0
