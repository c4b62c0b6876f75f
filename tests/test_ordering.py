"""Fill-reducing orderings of the SLS layer (galahad_b200/csrc/ordering.cpp): approximate minimum degree and nested
dissection.  The reference takes its orderings from MC68 / METIS, which its own tree only holds as stubs, so there is no
permutation to match: the tests check what any ordering must satisfy, and fill / flops against the
natural order and against the geometric nested dissection the benchmark problems come with."""
import ctypes as C

import numpy as np
import pytest

from galahad_b200 import problems as P
from galahad_b200._lib import Solver, as_ptr, default_options, lib
from galahad_b200.sls import SLS


def _order(prob, kind, base=0):
    L = lib()
    n = prob["n"]
    out = np.full(n, -7, np.int32)
    ptr = np.ascontiguousarray(prob["ptr"], np.int64) + base
    row = np.ascontiguousarray(prob["row"], np.int32) + base
    opt = default_options()
    opt.array_base = base
    rc = L.b200ldlt_order(n, as_ptr(ptr, C.c_int64), as_ptr(row, C.c_int), kind, as_ptr(out, C.c_int), C.byref(opt))
    return rc, out


def _fill(prob, order):
    s = Solver()
    assert s.analyse(prob["n"], prob["ptr"], prob["row"], order) >= 0
    out = (int(s.inform.num_factor), int(s.inform.num_flops))
    s.free()
    return out


PROBLEMS = [("laplace 3-D 16^3", lambda: P.laplace_problem(16)), ("KKT grid 10^3", lambda: P.kkt_grid_problem(10)),
            ("QPBAND 4000", lambda: P.qpband_problem(4000))]


@pytest.mark.parametrize("name,make", PROBLEMS)
@pytest.mark.parametrize("kind", [1, 3])
def test_orderings_are_permutations_and_reduce_fill(name, make, kind):
    prob = make()
    n = prob["n"]
    rc, o = _order(prob, kind)
    assert rc == 0 and sorted(o.tolist()) == list(range(n))
    rc1, o1 = _order(prob, kind, base=1)                  # Fortran indexing: the same order, shifted
    assert rc1 == 0 and np.array_equal(o1, o + 1)
    nf, fl = _fill(prob, o)
    nf_nat, fl_nat = _fill(prob, np.arange(n, dtype=np.int32))
    nf_geo, fl_geo = _fill(prob, prob["order"])
    assert nf < nf_nat and fl < fl_nat
    # within a small factor of the hand-made geometric nested dissection of these grid problems
    assert nf <= 1.5 * nf_geo and fl <= 2.5 * fl_geo, (name, kind, nf, nf_geo, fl, fl_geo)


def test_ordering_edge_cases():
    L = lib()
    opt = default_options()
    # empty matrix, diagonal matrix, one dense row (ordered last by the dense-row rule), two components
    assert L.b200ldlt_order(0, None, None, 1, None, C.byref(opt)) == 0
    n = 50
    diag = dict(n=n, ptr=np.arange(n + 1, dtype=np.int64), row=np.arange(n, dtype=np.int32))
    for kind in (1, 3):
        rc, o = _order(diag, kind)
        assert rc == 0 and sorted(o.tolist()) == list(range(n))
    n = 400                                                # arrow: variable 0 coupled to everybody
    ptr = np.concatenate([[0], np.arange(n, 2 * n)]).astype(np.int64)
    row = np.concatenate([np.arange(n), np.arange(1, n)]).astype(np.int32)
    arrow = dict(n=n, ptr=ptr, row=row)
    for kind in (1, 3):
        rc, o = _order(arrow, kind)
        assert rc == 0 and sorted(o.tolist()) == list(range(n))
        assert o[0] >= n - 2, (kind, o[0])                 # no fill: the hub goes last (or next to last)
    rc, o = _order(arrow, 2)
    assert rc == -98                                       # kinds other than 1 and 3 are not implemented


def test_sls_analyse_without_perm_uses_the_orderings():
    """control%ordering with no PERM: < 0 (the solver's own, the default -1) and 3 give nested dissection, 0 / 1 / 2 minimum
    degree, 4 / 5 / 6 GALAHAD_unavailable_option, -2 and below the reference's error codes; the order the solver reports is a permutation with less fill than the
    identity (sls.f90:2262-2268, 2395-2412)."""
    prob = P.laplace_problem(12)
    n = prob["n"]
    colidx = np.repeat(np.arange(n), np.diff(prob["ptr"]))
    nf_nat, _ = _fill(prob, np.arange(n, dtype=np.int32))
    seen = {}
    for oc in (-1, 0, 1, 2, 3, 7):
        s = SLS()
        s.control.ordering = oc
        assert s.analyse_matrix(n, "coordinate", prob["row"].size, prob["row"], colidx) == 0, oc
        inf = s.information()
        assert 0 < inf.entries_in_factors < nf_nat, (oc, inf.entries_in_factors, nf_nat)
        seen[oc] = inf.entries_in_factors
        s.terminate()
    assert seen[-1] == seen[3] and seen[0] == seen[1] == seen[2] == seen[7]
    # -2 = SSIDS's matching-based order, for which SLS_analyse has no values to pass: SSIDS_ERROR_VAL -> error_restrictions;
    # below that SSIDS_ERROR_ORDER -> error_inertia (the table of SLS_copy_inform_from_ssids, sls.f90:1766-1773)
    for oc, want in ((4, -29), (5, -29), (6, -29), (-2, -3), (-3, -20)):
        s = SLS()
        s.control.ordering = oc
        assert s.analyse_matrix(n, "coordinate", prob["row"].size, prob["row"], colidx) == want, oc
        s.terminate()
