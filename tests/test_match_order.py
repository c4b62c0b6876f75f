"""Matching-based ordering (SSIDS options%ordering = 2): the product's C++ match_order
(csrc/scaling.cpp, through the C ABI) against the independent Python restatement of the reference's Fortran
(oracle/scaling_oracle.py: match_order_metis / mo_match / mo_split of src/spral/match_order.f90 and expand_matrix of
src/ssids/anal.f90).  Bit for bit: the matching split into singletons and pairs, the compressed pattern, the scaling
and -- given the same order of the compressed graph, which the reference takes from METIS and this library from its own
nested dissection -- the pivot order.  Then the properties that define the result, at sizes the Python oracle is not run on."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from galahad_b200 import problems as P
from galahad_b200 import _lib
from oracle import scaling_oracle as HO
from test_scaling_oracle import small_cases, lower

I64, I32, F64 = C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_double)


def c_match_order(n, ptr, row, val, kind=3):
    L = _lib.lib()
    ptr = np.ascontiguousarray(ptr, np.int64)
    row = np.ascontiguousarray(row, np.int32)
    val = np.ascontiguousarray(val, np.float64)
    nz = int(ptr[n]) if n > 0 else 0
    order = np.zeros(max(n, 1), np.int32)
    scale = np.zeros(max(n, 1))
    pairs = np.zeros(max(n, 1), np.int32)
    ncomp = C.c_int(0)
    cptr = np.zeros(n + 2, np.int64)
    crow = np.zeros(2 * nz + 2, np.int32)
    flag = L.b200ldlt_match_order(n, ptr.ctypes.data_as(I64), row.ctypes.data_as(I32), val.ctypes.data_as(F64), kind,
                                  order.ctypes.data_as(I32), scale.ctypes.data_as(F64), pairs.ctypes.data_as(I32),
                                  C.byref(ncomp), cptr.ctypes.data_as(I64), crow.ctypes.data_as(I32))
    nc = ncomp.value
    return flag, order[:n], scale[:n], pairs[:n], nc, cptr[:nc + 1], crow[:int(cptr[nc]) if nc else 0]


def library_order(kind):
    """The order of the compressed graph as the library computes it (1-based lists in and out, as the oracle wants)."""
    def f(ncomp, ptr3, row3):
        L = _lib.lib()
        opt = _lib.default_options()
        opt.array_base = 0
        p = np.array([ptr3[i] - 1 for i in range(1, ncomp + 2)], np.int64)
        r = np.array([row3[k] - 1 for k in range(1, ptr3[ncomp + 1])], np.int32)
        o = np.zeros(max(ncomp, 1), np.int32)
        rc = L.b200ldlt_order(ncomp, p.ctypes.data_as(I64), r.ctypes.data_as(I32), kind, o.ctypes.data_as(I32), C.byref(opt))
        assert rc == 0
        return [0] + [int(v) + 1 for v in o[:ncomp]]
    return f


def extra_cases():
    rng = np.random.default_rng(5)
    out = []
    out.append(("one", 1, np.array([0, 1]), np.array([0]), np.array([3.0])))
    out.append(("diag", 5, np.arange(6), np.arange(5), np.array([1.0, -2.0, 3.0, 0.5, 8.0])))
    # an odd cycle: 3 x 3 with zero diagonal (matching 1 -> 2 -> 3 -> 1 must split into a pair and a singleton)
    A = np.array([[0, 1.0, 2.0], [1.0, 0, 3.0], [2.0, 3.0, 0]])
    out.append(("odd_cycle",) + lower(A))
    # longer cycles: circulant band with zero diagonal
    for n in (7, 12):
        A = np.zeros((n, n))
        for i in range(n):
            A[i, (i + 1) % n] = A[(i + 1) % n, i] = 1.0 + 0.1 * i
            A[i, (i + 3) % n] = A[(i + 3) % n, i] = 0.01
        out.append((f"cycle_{n}",) + lower(A))
    # KKT with a zero (2,2) block: every constraint row has to be matched off the diagonal
    for m in (8, 20):
        n = 2 * m
        H = np.diag(rng.uniform(1e-3, 1, n))
        J = (rng.random((m, n)) < 0.2) * rng.uniform(-2, 2, (m, n))
        J[np.arange(m), np.arange(m)] = 1.0 + rng.random(m)
        K = np.block([[H, J.T], [J, np.zeros((m, m))]])
        out.append((f"kkt_zero22_{m}",) + lower(K))
    # an empty column (structurally singular), explicit zero entries stored in the pattern
    A = np.diag([1.0, 2.0, 0.0, 4.0, 5.0, 6.0])
    A[4, 1] = A[1, 4] = 7.0
    A[5, 0] = A[0, 5] = 0.5
    n, ptr, row, val = lower(A + np.eye(6) * 1e-300)    # keep the diagonal of column 2 in the pattern ...
    val = np.where(np.abs(val) <= 1e-299, 0.0, val - 0.0)
    A2 = A.copy()
    out.append(("empty_col_explicit_zero", n, ptr, row, np.array([A2[r, c] for c in range(n) for r in row[ptr[c]:ptr[c + 1]]])))
    return out


CASES = small_cases() + extra_cases()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("kind", [3, 1])
def test_match_order_against_restatement(case, kind):
    name, n, ptr, row, val = case
    if kind == 1 and n > 150:
        pytest.skip("one ordering kind is enough for the larger cases")
    flag, order, scale, pairs, nc, cptr, crow = c_match_order(n, ptr, row, val, kind)
    fo, oo, so, po, (nco, ptr3, row3) = HO.match_order(n, ptr, row, val, library_order(kind))
    assert flag == fo
    assert list(pairs) == po
    assert nc == nco
    assert list(cptr) == [ptr3[i] - 1 for i in range(1, nco + 2)]
    assert list(crow) == [row3[k] - 1 for k in range(1, ptr3[nco + 1])]
    assert np.array_equal(scale.view(np.int64), np.array(so, np.float64).view(np.int64)), "scaling differs in the last bits"
    assert list(order) == oo
    check_properties(n, ptr, row, val, flag, order, scale, pairs)


def check_properties(n, ptr, row, val, flag, order, scale, pairs):
    assert sorted(order) == list(range(n))
    col = np.repeat(np.arange(n), np.diff(ptr))
    a = np.abs(val) * scale[row] * scale[col]
    assert (a <= 1 + 1e-9).all()
    lookup = {(int(r), int(c)): v for r, c, v in zip(row, col, a)}
    for i in range(n):
        j = int(pairs[i])
        if j > 0:
            j -= 1
            assert pairs[j] == i + 1
            assert abs(order[i] - order[j]) == 1, "a matched pair takes consecutive pivots"
            key = (max(i, j), min(i, j))
            assert key in lookup
            if flag == 0:
                assert abs(lookup[key] - 1) < 1e-9, "matched entries scale to 1"
        elif j == -2:
            assert flag == 1 and scale[i] == 0.0
        else:
            assert j == -1
    if flag == 0:
        assert (pairs != -2).all()


def test_larger_problems_properties():
    for prob in (P.kkt_grid_problem(14), P.qpband_problem(4000), P.laplace_problem(16)):
        n, ptr, row, val = prob["n"], prob["ptr"], prob["row"], prob["val"]
        flag, order, scale, pairs, nc, cptr, crow = c_match_order(n, ptr, row, val)
        assert flag == 0
        check_properties(n, ptr, row, val, flag, order, scale, pairs)
        assert nc == n - int((pairs > 0).sum()) // 2


def test_kkt_zero_block_pairs_every_constraint():
    rng = np.random.default_rng(3)
    g = 8
    nv = g ** 3
    m = nv // 2
    H = sp.diags(rng.uniform(1e-4, 1e-2, nv))
    rows = np.repeat(np.arange(m), 2)
    cols = np.stack([np.arange(m), np.arange(m) + m], 1).ravel()
    J = sp.csr_matrix((np.ones(2 * m), (rows, cols)), shape=(m, nv))
    K = sp.bmat([[H, J.T], [J, None]], format="csc")
    n, ptr, row, val = P.lower_csc(K)
    flag, order, scale, pairs, nc, cptr, crow = c_match_order(n, ptr, row, val)
    assert flag == 0
    assert (pairs[nv:] > 0).all(), "a zero diagonal can only be matched off the diagonal"
    assert (pairs[nv:] <= nv).all()
    check_properties(n, ptr, row, val, flag, order, scale, pairs)


def test_symbols_and_argument_checks():
    L = _lib.lib()
    assert hasattr(L, "b200ldlt_match_order") and hasattr(L, "b200ldlt_analyse_matched")
    assert L.b200ldlt_match_order(-1, None, None, None, 3, None, None, None, None, None, None) < 0
    s = _lib.Solver()
    ptr = np.array([0, 1, 2], np.int64)
    row = np.array([0, 1], np.int32)
    order = np.zeros(2, np.int32)
    rc = L.b200ldlt_analyse_matched(C.byref(s.akeep), 2, ptr.ctypes.data_as(I64), row.ctypes.data_as(I32), None,
                                    order.ctypes.data_as(I32), C.byref(s.opt), C.byref(s.inform))
    assert rc == -9          # SSIDS_ERROR_VAL (ssids.f90:230-238)


def test_analyse_matched_end_to_end_on_the_host():
    """b200ldlt_analyse_matched = ssids_analyse with options%ordering = 2 (ssids.f90:321-358): the symbolic analysis runs on
    the order match_order gives (same factor size and flops as analysing with that order), a structurally singular matrix
    raises the warning SSIDS_WARNING_ANAL_SINGULAR (6) and still analyses."""
    name, n, ptr, row, val = [c for c in CASES if c[0] == "kkt_zero22_20"][0]
    s = _lib.Solver()
    flag = s.analyse_matched(n, ptr, row, val)
    assert flag == 0 and s.inform.flag == 0
    f2, order, scale, pairs, nc, cptr, crow = c_match_order(n, ptr, row, val)
    # analyse relabels the order by its postorder of the elimination tree (as ssids_analyse does with the order
    # match_order_metis hands it): a permutation with the fill of the matching-based order
    assert sorted(s.order) == list(range(n))
    assert s.inform.num_factor > 0 and s.inform.num_sup > 0 and s.inform.matrix_rank == n
    s2 = _lib.Solver()
    assert s2.analyse(n, ptr, row, order) == 0
    assert (s2.inform.num_factor, s2.inform.num_flops) == (s.inform.num_factor, s.inform.num_flops)
    # structurally singular
    name, n, ptr, row, val = [c for c in CASES if c[0] == "singular_dense_col"][0]
    s3 = _lib.Solver()
    assert s3.analyse_matched(n, ptr, row, val) == 6
    assert sorted(s3.order) == list(range(n)) and s3.inform.num_factor > 0
