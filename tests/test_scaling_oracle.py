"""Matching-based scaling (SSIDS options%scaling = 1): the product's C++ Hungarian scaling
(csrc/scaling.cpp, through the C ABI) against the independent Python restatement of the reference's Fortran
(oracle/scaling_oracle.py): scaling factors, matching and its cardinality, entry by entry and to the bit (both follow
scaling.f90's scan orders and use the C library's log / exp); then the properties that define the scaling -- every
|s_i a_ij s_j| <= 1, = 1 on the matching -- at sizes the Python oracle is not run on."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from galahad_b200 import problems as P
from galahad_b200 import _lib
from oracle import scaling_oracle as HO


def c_hungarian(n, ptr, row, val, scale_if_singular=True):
    L = _lib.lib()
    ptr = np.ascontiguousarray(ptr, np.int64)
    row = np.ascontiguousarray(row, np.int32)
    val = np.ascontiguousarray(val, np.float64)
    s = np.zeros(max(n, 1))
    match = np.zeros(max(n, 1), np.int32)
    matched = C.c_int(0)
    flag = L.b200ldlt_hungarian_scale(n, ptr.ctypes.data_as(C.POINTER(C.c_int64)), row.ctypes.data_as(C.POINTER(C.c_int)),
                                      val.ctypes.data_as(C.POINTER(C.c_double)), s.ctypes.data_as(C.POINTER(C.c_double)),
                                      match.ctypes.data_as(C.POINTER(C.c_int)), C.byref(matched), int(scale_if_singular))
    return flag, s[:n], match[:n], matched.value


def lower(A):
    n, ptr, row, val = P.lower_csc(sp.csc_matrix(A))
    return n, ptr, row, val


def small_cases():
    rng = np.random.default_rng(11)
    out = []
    for name, prob in (("kkt7_6", P.kkt_grid_problem(6)), ("lap7_6", P.laplace_problem(6)), ("qpband_300", P.qpband_problem(300))):
        out.append((name, prob["n"], prob["ptr"], prob["row"], prob["val"]))
    prob = P.kkt_grid_problem(6)
    col = np.repeat(np.arange(prob["n"]), np.diff(prob["ptr"]))
    d = 10.0 ** rng.uniform(-6, 6, prob["n"])
    out.append(("kkt7_6_misscaled", prob["n"], prob["ptr"], prob["row"], prob["val"] * d[prob["row"]] * d[col]))
    for t in range(4):
        n = 120
        M = sp.random(n, n, density=0.03, random_state=int(rng.integers(1 << 30)), format="csc", data_rvs=lambda k: 10.0 ** rng.uniform(-4, 4, k))
        M = M + M.T
        if t % 2 == 0:
            M = M + sp.diags(10.0 ** rng.uniform(-3, 3, n))
        out.append((f"sprand_{t}",) + lower(M.toarray()))
    # zero diagonal saddle point (matching must leave the diagonal), with explicit zeros among the entries
    n = 40
    A = np.zeros((n, n))
    B = rng.uniform(-1, 1, (n // 2, n // 2)) * (rng.random((n // 2, n // 2)) < 0.3) + np.eye(n // 2) * 2
    A[n // 2:, :n // 2] = B
    A = A + A.T
    out.append(("saddle_zero_diag",) + lower(A))
    # structurally singular: two empty-ish rows, and a dense column (more than m / 10 entries with m > 50)
    n = 60
    A = np.diag(rng.uniform(1, 2, n))
    A[:, 7] = rng.uniform(0.1, 1, n)
    A[7, :] = A[:, 7]
    A[20, :] = 0; A[:, 20] = 0
    A[41, :] = 0; A[:, 41] = 0
    A[41, 3] = A[3, 41] = 0.5
    A[20, 3] = A[3, 20] = 0.25        # rows 20 and 41 both only reach column 3: rank deficient
    out.append(("singular_dense_col",) + lower(A))
    return out


@pytest.mark.parametrize("case", small_cases(), ids=[c[0] for c in small_cases()])
def test_product_matches_scaling_oracle(case):
    name, n, ptr, row, val = case
    for sis in (True, False):
        fo, so, mo, ko = HO.hungarian_scale_sym(n, ptr, row, val, scale_if_singular=sis)
        fc, sc, mc, kc = c_hungarian(n, ptr, row, val, sis)
        assert fc == fo, name
        assert kc == ko and np.array_equal(mc, np.asarray(mo, np.int32)), name
        assert np.array_equal(sc, np.asarray(so)), (name, float(np.abs(sc - np.asarray(so)).max()))
    if name.startswith("singular"):
        assert fo == -2 and np.all(sc == 1.0)           # scale_if_singular = False: identity scaling, error flag
        f1, s1, m1, k1 = c_hungarian(n, ptr, row, val, True)
        assert f1 == 1 and k1 < n and np.all(np.isfinite(s1)) and np.all(s1 > 0)


@pytest.mark.parametrize("name", ["kkt7_16_misscaled", "lap27_12", "sprand_3000"])
def test_scaling_properties(name):
    rng = np.random.default_rng(3)
    if name == "kkt7_16_misscaled":
        prob = P.kkt_grid_problem(16)
        n, ptr, row = prob["n"], prob["ptr"], prob["row"]
        col = np.repeat(np.arange(n), np.diff(ptr))
        d = 10.0 ** rng.uniform(-8, 8, n)
        val = prob["val"] * d[row] * d[col]
    elif name == "lap27_12":
        prob = P.laplace_problem(12, stencil=27)
        n, ptr, row, val = prob["n"], prob["ptr"], prob["row"], prob["val"]
    else:
        n = 3000
        M = sp.random(n, n, density=0.002, random_state=5, format="csc", data_rvs=lambda k: 10.0 ** rng.uniform(-5, 5, k))
        M = M + M.T + sp.diags(10.0 ** rng.uniform(-5, 5, n))
        n, ptr, row, val = lower(M)
    flag, s, match, matched = c_hungarian(n, ptr, row, val)
    assert flag == 0 and matched == n and np.all(s > 0)
    assert sorted(match.tolist()) == list(range(1, n + 1))                 # a permutation
    col = np.repeat(np.arange(n), np.diff(ptr))
    scaled = np.abs(val) * s[row] * s[col]
    assert scaled.max() <= 1.0 + 1e-12                                      # dual feasibility
    # complementary slackness on the matching: entry (i, match[i]) of the full matrix has modulus 1 after scaling
    full = sp.csc_matrix((np.abs(val), (row, col)), shape=(n, n))
    full = full + sp.tril(full, -1).T
    full = full.tocsr()
    S = sp.diags(s)
    scaled_full = (S @ full @ S).tocsr()
    on_match = np.asarray(scaled_full[np.arange(n), match - 1]).ravel()
    assert np.abs(on_match - 1.0).max() <= 1e-10


def c_auction(n, ptr, row, val):
    L = _lib.lib()
    ptr = np.ascontiguousarray(ptr, np.int64)
    row = np.ascontiguousarray(row, np.int32)
    val = np.ascontiguousarray(val, np.float64)
    s = np.zeros(max(n, 1))
    match = np.zeros(max(n, 1), np.int32)
    matched, its = C.c_int(0), C.c_int(0)
    flag = L.b200ldlt_auction_scale(n, ptr.ctypes.data_as(C.POINTER(C.c_int64)), row.ctypes.data_as(C.POINTER(C.c_int)),
                                    val.ctypes.data_as(C.POINTER(C.c_double)), s.ctypes.data_as(C.POINTER(C.c_double)),
                                    match.ctypes.data_as(C.POINTER(C.c_int)), C.byref(matched), C.byref(its))
    return flag, s[:n], match[:n], matched.value, its.value


@pytest.mark.parametrize("case", small_cases(), ids=[c[0] for c in small_cases()])
def test_auction_matches_scaling_oracle(case):
    """options%scaling = 2: SPRAL's auction algorithm (scaling.f90:269-309, 1351-1609), C++ against the Python restatement:
    scaling, matching, cardinality and sweep count to the bit; the scaling is positive and finite, and the scaled matrix
    has no entry above e (the auction's duals are feasible up to eps <= 1 per row and column pair)."""
    name, n, ptr, row, val = case
    so, mo, ko, io = HO.auction_scale_sym(n, ptr, row, val)
    fc, sc, mc, kc, ic = c_auction(n, ptr, row, val)
    assert fc == 0 and kc == ko and ic == io, name
    assert np.array_equal(mc, np.asarray(mo, np.int32)), name
    assert np.array_equal(sc, np.asarray(so)), (name, float(np.abs(sc - np.asarray(so)).max()))
    assert np.all(np.isfinite(sc)) and np.all(sc > 0)
    if not name.startswith("singular"):
        assert kc >= 0.9 * n
        col = np.repeat(np.arange(n), np.diff(ptr))
        assert (np.abs(val) * sc[row] * sc[col]).max() <= np.e * (1 + 1e-12)


def test_scaling_feeds_the_engine_options():
    # options%scaling = 2, 3 stay unimplemented (-98): checked without a GPU through the argument validation order
    L = _lib.lib()
    assert hasattr(L, "b200ldlt_hungarian_scale")
