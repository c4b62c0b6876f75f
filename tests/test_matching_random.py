"""Randomised sweep of the matching-based routines: the product's C++ (csrc/scaling.cpp, through the C ABI) against the
Python restatements of the reference's Fortran (oracle/scaling_oracle.py) on a few hundred small random symmetric
matrices -- dense and sparse, zero diagonals, explicit zeros, empty and duplicate-pattern rows, structurally singular,
badly scaled.  The fixed cases of test_scaling_oracle.py / test_match_order.py pin the interesting shapes; this one is
after the rare branches (heap deletions, ties, second matchings of the singular path).  Everything to the bit."""
import ctypes as C

import numpy as np
import pytest

from galahad_b200 import _lib
from oracle import scaling_oracle as HO
from test_match_order import c_match_order, library_order
from test_scaling_oracle import c_auction, c_hungarian


def random_lower(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 36))
    kind = seed % 6
    dens = [0.08, 0.2, 0.5, 1.0, 0.15, 0.3][kind]
    A = np.zeros((n, n))
    mask = np.tril(rng.random((n, n)) < dens, -1)
    mag = 10.0 ** rng.uniform(-6, 6, (n, n)) if kind in (1, 4) else rng.uniform(0.1, 2.0, (n, n))
    A[mask] = (mag * rng.choice([-1.0, 1.0], (n, n)))[mask]
    diag = rng.uniform(0.1, 2.0, n) * rng.choice([-1.0, 1.0], n)
    if kind in (2, 4):
        diag[rng.random(n) < 0.6] = 0.0                  # zero diagonals: the matching has to leave the diagonal
    if kind == 5 and n > 3:                              # rank deficient: rows that only see one common column
        k = int(rng.integers(0, n))
        for i in rng.choice(n, size=min(3, n), replace=False):
            A[i, :] = 0.0; A[:, i] = 0.0; diag[i] = 0.0
            if i != k:
                A[max(i, k), min(i, k)] = 0.5
    if kind == 3:                                        # ties: small integer entries
        A = np.sign(A) * np.ceil(np.abs(A))
        diag = np.sign(diag) * np.ceil(np.abs(diag))
    A[np.arange(n), np.arange(n)] = diag
    ptr = [0]
    row, val = [], []
    for c in range(n):
        for r in range(c, n):
            explicit_zero = (r == c) or (rng.random() < 0.03)     # the diagonal is always in the pattern (SLS's CSR)
            if A[r, c] != 0.0 or explicit_zero:
                row.append(r); val.append(A[r, c])
        ptr.append(len(row))
    return n, np.array(ptr, np.int64), np.array(row, np.int32), np.array(val, np.float64)


SEEDS = list(range(240))


@pytest.mark.parametrize("chunk", range(8))
def test_random_matrices_to_the_bit(chunk):
    for seed in SEEDS[chunk::8]:
        n, ptr, row, val = random_lower(seed)
        for sis in (True, False):
            fo, so, mo, ko = HO.hungarian_scale_sym(n, ptr, row, val, scale_if_singular=sis)
            fc, sc, mc, kc = c_hungarian(n, ptr, row, val, sis)
            assert (fc, kc) == (fo, ko), seed
            assert np.array_equal(mc, np.asarray(mo, np.int32)), seed
            assert np.array_equal(sc.view(np.int64), np.asarray(so, np.float64).view(np.int64)), seed
        so, mo, ko, io = HO.auction_scale_sym(n, ptr, row, val)
        fc, sc, mc, kc, ic = c_auction(n, ptr, row, val)
        assert (fc, kc, ic) == (0, ko, io), seed
        assert np.array_equal(mc, np.asarray(mo, np.int32)), seed
        assert np.array_equal(sc.view(np.int64), np.asarray(so, np.float64).view(np.int64)), seed
        flag, order, scale, pairs, nc, cptr, crow = c_match_order(n, ptr, row, val, 1)
        fo, oo, so, po, (nco, ptr3, row3) = HO.match_order(n, ptr, row, val, library_order(1))
        assert (flag, nc) == (fo, nco), seed
        assert list(pairs) == po and list(order) == oo, seed
        assert list(cptr) == [ptr3[i] - 1 for i in range(1, nco + 2)], seed
        assert list(crow) == [row3[k] - 1 for k in range(1, ptr3[nco + 1])], seed
        assert np.array_equal(scale.view(np.int64), np.asarray(so, np.float64).view(np.int64)), seed
        assert sorted(order) == list(range(n)), seed
