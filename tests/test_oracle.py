"""Pins the oracle: (i) the reference's own known-answer systems and bcsstk01, (ii) the committed
outputs of the reference's C++ engine (tests/golden/ref_runs.json), and (iii) -- when oracle/_ref is
present -- a live run of that engine on the same seeded inputs."""
import numpy as np
import pytest

from galahad_b200 import problems as P
from galahad_b200._lib import Solver
from oracle import refssids
from oracle.refssids import Oracle, RefSSIDS, permuted_solve
from util import bcsstk01_problem, golden, golden_cases, kat_problem, rhs_for

EPS = np.finfo(float).eps


def analyse(prob):
    s = Solver()
    assert s.analyse(prob["n"], prob["ptr"], prob["row"], prob["order"]) >= 0
    return s, s.symbolic()


@pytest.mark.parametrize("which", ["indef", "spd"])
@pytest.mark.parametrize("order", ["identity", "reverse"])
def test_kat_5x5(which, order):
    prob, rhs, sol = kat_problem(which, order)
    s, sym = analyse(prob)
    for eng in ([Oracle(sym, 5)] + ([RefSSIDS(sym, 5)] if refssids.have_ref() else [])):
        st = eng.factor(prob["val"], False)
        assert st["flag"] == 0
        x = permuted_solve(eng, s.order, rhs)
        assert np.abs(x - sol).max() <= EPS ** 0.5          # slst.f90:261-262
    if which == "spd":
        orc = Oracle(sym, 5)
        assert orc.factor(prob["val"], True)["flag"] == 0   # Cholesky path
        assert np.abs(permuted_solve(orc, s.order, rhs) - sol).max() <= EPS ** 0.5


@pytest.mark.parametrize("order", ["identity", "reverse"])
def test_bcsstk01(order):
    prob = bcsstk01_problem(order)
    s, sym = analyse(prob)
    A = P.full_from_lower(prob["n"], prob["ptr"], prob["row"], prob["val"])
    b = A @ np.ones(prob["n"])                              # runsls_rb.f90:126-137
    orc = Oracle(sym, prob["n"])
    st = orc.factor(prob["val"], False)
    assert (st["flag"], st["num_neg"], st["num_zero"]) == (0, 0, 0)
    x = permuted_solve(orc, s.order, b)
    assert P.backward_error(prob, x, b) <= 1e-14
    assert np.abs(x - 1).max() <= 1e-9                      # cond ~ 8.8e5


def test_against_committed_reference_runs():
    gold = {r["name"]: r for r in golden("ref_runs.json")["runs"]}
    for name, prob, posdef in golden_cases():
        g = gold[name]
        s, sym = analyse(prob)
        assert (sym["nnodes"], s.inform.num_factor, s.inform.num_flops) == (g["num_sup"], g["num_factor"], g["num_flops"])
        orc = Oracle(sym, prob["n"])
        st = orc.factor(prob["val"], posdef)
        assert st["flag"] == g["flag"], name
        assert st["num_neg"] == g["num_neg"], name          # inertia: bit exact
        assert st["num_zero"] == g["num_zero"], name
        xs, b = rhs_for(prob)
        x = permuted_solve(orc, s.order, b)
        assert P.backward_error(prob, x, b) <= max(1e-14, 10 * g["berr"]), name
        scale = max(1.0, np.abs(x).max())
        assert np.abs(x[:8] - np.array(g["x_head"])).max() <= 1e-8 * scale, name


@pytest.mark.skipif(not refssids.have_ref(), reason="oracle/_ref not built")
def test_live_reference_engine():
    for name, prob, posdef in golden_cases()[:6] + golden_cases()[-3:]:
        s, sym = analyse(prob)
        ref, orc = RefSSIDS(sym, prob["n"]), Oracle(sym, prob["n"])
        a, b_ = ref.factor(prob["val"], posdef), orc.factor(prob["val"], posdef)
        assert (a["flag"], a["num_neg"], a["num_zero"]) == (b_["flag"], b_["num_neg"], b_["num_zero"]), name
        xs, b = rhs_for(prob)
        xr, xo = permuted_solve(ref, s.order, b), permuted_solve(orc, s.order, b)
        assert np.abs(xr - xo).max() <= 1e-8 * max(1.0, np.abs(xr).max()), name
        # partial solves (jobs 1..4, datatypes.f90:62-66) compose to the full solve
        if not posdef:
            y = permuted_solve(orc, s.order, b, job=1)
            z = permuted_solve(orc, s.order, y, job=4)
            assert np.abs(z - xo).max() <= 1e-12 * max(1.0, np.abs(xo).max())


def test_singular_matrix_is_flagged():
    import scipy.sparse as sp
    A = sp.csc_matrix(np.array([[1.0, 1.0, 0], [1.0, 1.0, 0], [0, 0, 2.0]]))
    n, ptr, row, val = P.lower_csc(A)
    prob = dict(n=n, ptr=ptr, row=row, val=val, order=np.arange(3, dtype=np.int32))
    s, sym = analyse(prob)
    st = Oracle(sym, 3, small=1e-14).factor(val, False)
    assert st["flag"] == 7 and st["num_zero"] == 1          # SSIDS_WARNING_FACT_SINGULAR
    st = Oracle(sym, 3, small=1e-14, action=False).factor(val, False)
    assert st["flag"] == -5                                 # SSIDS_ERROR_SINGULAR


def _equilib_numpy(n, ptr, row, val, max_iterations=10, tol=np.float32(1e-8)):
    """Independent restatement of inf_norm_equilib_sym (src/spral/scaling.f90:480-521), vectorised: the
    per-row maxima are order-independent, every other operation is elementwise, so the bits must agree."""
    col = np.repeat(np.arange(n), np.diff(ptr))
    s = np.ones(n)
    it = 0
    for it in range(1, max_iterations + 1):
        v = np.abs(s[row] * val * s[col])
        m = np.zeros(n)
        np.maximum.at(m, row, v)
        np.maximum.at(m, col, v)
        nzr = m > 0
        s[nzr] = s[nzr] / np.sqrt(m[nzr])
        if np.abs(1 - m).max() < float(tol):
            break
    else:
        it = max_iterations + 1
    return s, it - 1


def test_equilibration_scaling_matches_restatement():
    """options%scaling = 4: the C++ restatement behind b200ldlt_factor_scaled against the numpy one, bit for
    bit, on badly scaled KKT matrices; and the property the algorithm is for: every row of |S A S| has
    infinity norm 1 to the tolerance once it has converged."""
    import ctypes as C
    from galahad_b200 import _lib
    from galahad_b200 import problems as P
    L = _lib.lib()
    rng = np.random.default_rng(5)
    for prob in (P.kkt_grid_problem(6, dlo=1e-10), P.random_sym_problem(300, 0.02, "indef", seed=9)):
        n, ptr, row = prob["n"], np.asarray(prob["ptr"], np.int64), np.asarray(prob["row"], np.int32)
        d = 10.0 ** rng.uniform(-6, 6, n)                      # destroy the scaling symmetrically
        col = np.repeat(np.arange(n), np.diff(ptr))
        val = prob["val"] * d[row] * d[col]
        s = np.zeros(n)
        it = C.c_int()
        assert L.b200ldlt_equilib_scale(n, ptr.ctypes.data_as(C.POINTER(C.c_int64)), row.ctypes.data_as(C.POINTER(C.c_int)),
                                        val.ctypes.data_as(C.POINTER(C.c_double)), s.ctypes.data_as(C.POINTER(C.c_double)),
                                        C.byref(it)) == 0
        s_np, it_np = _equilib_numpy(n, ptr, row, val)
        assert it.value == it_np and np.array_equal(s, s_np)
        # more iterations: converged scaling equilibrates the rows
        s2, it2 = _equilib_numpy(n, ptr, row, val, max_iterations=200)
        v = np.abs(s2[row] * val * s2[col])
        m = np.zeros(n)
        np.maximum.at(m, row, v)
        np.maximum.at(m, col, v)
        assert it2 < 200 and np.abs(1 - m).max() < 1e-6


@pytest.mark.skipif(not refssids.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("chunk", range(3))
def test_restatement_against_live_reference_on_random_matrices(chunk):
    """The pin beyond the committed runs: 36 random sparse matrices (SPD, indefinite with tiny diagonals, KKT; random pivot
    orders; nemin 1 / 8 / 32 so that fronts come in all shapes) through the reference's own engine and through the C
    restatement: flags, inertia and zero pivots exact, solutions equal to rounding.  (The number of 2 x 2 pivots is not
    compared: the reference sends ragged blocks -- fronts narrower than 32 columns, the last columns of a front -- through
    its TPP kernel, which takes 2 x 2 pivots even on SPD matrices, where the restatement keeps block pivoting; the pivot
    sequence is not part of the result contract -- DESIGN.md section 5, known divergences.)"""
    for t in range(chunk, 36, 3):
        kind = ("spd", "indef", "kkt")[t % 3]
        n = 40 + 29 * (t % 9)
        prob = P.random_sym_problem(n, [0.02, 0.05, 0.12][(t // 3) % 3], kind, seed=9000 + t)
        s = Solver()
        s.opt.nemin = [1, 8, 32][(t // 9) % 3]
        assert s.analyse(prob["n"], prob["ptr"], prob["row"], prob["order"]) >= 0
        sym = s.symbolic()
        ref, orc = RefSSIDS(sym, prob["n"]), Oracle(sym, prob["n"])
        a, b_ = ref.factor(prob["val"], False), orc.factor(prob["val"], False)
        for k in ("flag", "num_neg", "num_zero"):
            assert a[k] == b_[k], (t, k, a[k], b_[k])
        xs, b = rhs_for(prob, seed=t)
        xr, xo = permuted_solve(ref, s.order, b), permuted_solve(orc, s.order, b)
        assert np.abs(xr - xo).max() <= 1e-7 * max(1.0, np.abs(xr).max()), (t, np.abs(xr - xo).max())
