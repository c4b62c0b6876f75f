"""GPU parity tests: the CUDA engine, called through the C ABI, against the oracle on the same
seeded inputs.  Integer results (inertia, delays reported by both as they use the same schedule)
are compared exactly; solutions by scaled backward error <= 1e-14 and agreement with the oracle's
solution to 1e-10 x a condition allowance."""
import ctypes as C

import numpy as np
import pytest

from galahad_b200 import problems as P
from galahad_b200 import _lib
from galahad_b200._lib import Solver
from oracle.refssids import Oracle, permuted_solve
from util import bcsstk01_problem, golden, golden_cases, kat_problem, rhs_for

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


def gpu_factor(prob, posdef=False, **opts):
    s = Solver(**opts)
    assert s.analyse(prob["n"], prob["ptr"], prob["row"], prob["order"]) >= 0
    flag = s.factor(prob["val"], posdef)
    assert flag != -51, "CUDA engine unavailable: the product path has no CPU fallback"
    return s, flag


@pytest.mark.parametrize("which", ["indef", "spd"])
@pytest.mark.parametrize("order", ["identity", "reverse"])
def test_kat_5x5(which, order):
    prob, rhs, sol = kat_problem(which, order)
    for posdef in ([False, True] if which == "spd" else [False]):
        s, flag = gpu_factor(prob, posdef)
        assert flag == 0
        f, x = s.solve(rhs)
        assert f == 0 and np.abs(x - sol).max() <= EPS ** 0.5
        f, x2 = s.solve(np.stack([rhs, 2 * rhs], axis=1))        # 2 RHS (slst.f90)
        assert np.abs(x2[:, 0] - sol).max() <= EPS ** 0.5 and np.abs(x2[:, 1] - 2 * sol).max() <= EPS ** 0.5


def test_bcsstk01():
    for order in ("identity", "reverse"):
        prob = bcsstk01_problem(order)
        s, flag = gpu_factor(prob)
        assert flag == 0 and s.inform.num_neg == 0 and s.inform.matrix_rank == 48
        b = P.full_from_lower(prob["n"], prob["ptr"], prob["row"], prob["val"]) @ np.ones(48)
        f, x = s.solve(b)
        assert P.backward_error(prob, x, b) <= 1e-14 and np.abs(x - 1).max() <= 1e-9


def test_against_oracle_and_reference_runs():
    gold = {r["name"]: r for r in golden("ref_runs.json")["runs"]}
    for name, prob, posdef in golden_cases():
        g = gold[name]
        s, flag = gpu_factor(prob, posdef)
        sym = s.numeric_tree()    # same fronts as the engine (chains fused), so results are comparable front by front
        orc = Oracle(sym, prob["n"])
        so = orc.factor(prob["val"], posdef)
        assert flag == so["flag"] == g["flag"], name
        assert s.inform.num_neg == so["num_neg"] == g["num_neg"], name       # bit exact
        xs, b = rhs_for(prob)
        f, x = s.solve(b)
        xo = permuted_solve(orc, s.order, b)
        assert f == 0
        assert P.backward_error(prob, x, b) <= max(1e-14, 10 * g["berr"]), name
        assert np.abs(x - xo).max() <= 1e-10 * 1e4 * max(1.0, np.abs(xo).max()), name


@pytest.mark.parametrize("chunk", range(3))
def test_random_sweep_against_oracle(chunk):
    """Beyond the committed golden runs: 30 random sparse matrices (SPD, indefinite with small diagonals -- 2 x 2 pivots,
    failed and delayed columns -- and KKT with a tiny (2,2) block) under random pivot orders, the C restatement of the
    reference engine factorising the same fronts: inertia exact, both solutions backward stable, and equal up to the
    conditioning the oracle itself shows."""
    for t in range(chunk, 30, 3):
        kind = ("spd", "indef", "kkt")[t % 3]
        n = 60 + 37 * (t % 11)
        prob = P.random_sym_problem(n, [0.01, 0.03, 0.08][(t // 3) % 3], kind, seed=5000 + t)
        s, flag = gpu_factor(prob, posdef=False)
        sym = s.numeric_tree()
        orc = Oracle(sym, prob["n"])
        so = orc.factor(prob["val"], False)
        assert flag == so["flag"] == 0, (t, flag, so["flag"])
        assert s.inform.num_neg == so["num_neg"], t
        if kind == "spd":
            assert s.inform.num_neg == 0 and s.inform.num_delay == 0
        xs, b = rhs_for(prob, seed=t)
        f, x = s.solve(b)
        xo = permuted_solve(orc, s.order, b)
        assert f == 0
        be, beo = P.backward_error(prob, x, b), P.backward_error(prob, xo, b)
        assert be <= max(1e-14, 10 * beo), (t, be, beo)
        assert np.abs(x - xo).max() <= 1e-6 * max(1.0, np.abs(xo).max()), (t, np.abs(x - xo).max())


def test_partial_solves_and_enquire():
    prob = P.kkt_grid_problem(8)
    s, flag = gpu_factor(prob)
    xs, b = rhs_for(prob)
    f, x = s.solve(b)
    f, y = s.solve(b, job=1)
    f, z = s.solve(y, job=4)
    assert np.abs(z - x).max() <= 1e-12 * max(1, np.abs(x).max())
    f, y2 = s.solve(y, job=2)
    f, z2 = s.solve(y2, job=3)
    assert np.abs(z2 - x).max() <= 1e-12 * max(1, np.abs(x).max())
    f, piv, d = s.enquire_indef()
    assert f == 0 and sorted(np.abs(piv).tolist()) == list(range(prob["n"]))
    # inertia from D^{-1}: 1x1 signs + 2x2 determinants
    neg, k = 0, 0
    while k < prob["n"]:
        if d[k, 1] != 0.0:
            det = d[k, 0] * d[k + 1, 0] - d[k, 1] ** 2
            neg += 1 if det < 0 else (2 if d[k, 0] + d[k + 1, 0] < 0 else 0)
            k += 2
        else:
            neg += d[k, 0] < 0
            k += 1
    assert neg == s.inform.num_neg == prob["inertia"][1]
    # alter: scaling D^{-1} by 2 halves the solution
    assert s.alter(2 * d) == 0
    f, xh = s.solve(b)
    assert np.abs(xh - 2 * x).max() <= 1e-12 * max(1, np.abs(x).max())


def test_refactor_reuses_handles():
    prob = P.kkt_grid_problem(8)
    s, flag = gpu_factor(prob)
    for it in range(3):
        val = prob["val"].copy()
        val[prob["ptr"][:-1]] *= (1.0 + 0.1 * it)             # scale the diagonal, same pattern
        assert s.factor(val) >= 0
        p2 = dict(prob, val=val)
        xs, b = rhs_for(p2)
        f, x = s.solve(b)
        assert P.backward_error(p2, x, b) <= 1e-14


def test_singular_and_not_posdef():
    import scipy.sparse as sp
    A = sp.csc_matrix(np.array([[1.0, 1.0, 0], [1.0, 1.0, 0], [0, 0, 2.0]]))
    n, ptr, row, val = P.lower_csc(A)
    prob = dict(n=n, ptr=ptr, row=row, val=val, order=np.arange(3, dtype=np.int32))
    s, flag = gpu_factor(prob, small=1e-14)
    assert flag == 7 and s.inform.matrix_rank == 2
    s, flag = gpu_factor(prob, small=1e-14, action=0)
    assert flag == -5
    prob2, rhs, sol = kat_problem("indef")
    s, flag = gpu_factor(prob2, posdef=True)
    assert flag == -6                                         # SSIDS_ERROR_NOT_POS_DEF


def test_medium_grid_properties():
    """Size-independent properties at a size the oracle does not need to follow: exact inertia law
    of the KKT matrix (n, m, 0) and the backward-error bound."""
    for prob in (P.laplace_problem(24), P.kkt_grid_problem(16), P.qpband_problem(20000)):
        s, flag = gpu_factor(prob)
        assert flag == 0
        assert s.inform.num_neg == prob["inertia"][1]
        xs, b = rhs_for(prob)
        f, x = s.solve(b)
        assert P.backward_error(prob, x, b) <= 1e-14


def test_dense_fronts_with_failed_pivots():
    """One dense front per matrix: zero-diagonal / KKT-like matrices force 2x2 pivots, failed columns
    (a-posteriori test), shortened blocks, TPP and the retry rounds; orders above 1024 run under the
    look-ahead schedule (a failed pivot stalls the front until the synchronous retry rounds).  Inertia
    against the eigenvalues (exact), solutions by backward error."""
    for kind, n in (("zdiag", 142), ("kkt", 517), ("indef", 1300), ("zdiag", 2100), ("spd", 1100)):
        A, prob = P.dense_front_problem(n, kind, seed=n)
        s, flag = gpu_factor(prob)
        assert flag >= 0, (kind, n, flag)
        assert s.inform.num_neg == prob["inertia"][1], (kind, n)
        b = A @ np.ones(n)
        f, x = s.solve(b)
        be = np.abs(A @ x - b).max() / (np.abs(A).sum(1).max() * np.abs(x).max() + np.abs(b).max())
        assert f == 0 and be <= 1e-13, (kind, n, be)
        if kind == "spd":
            s2, flag2 = gpu_factor(prob, posdef=True)
            f, x2 = s2.solve(b)
            assert flag2 == 0 and np.abs(x2 - x).max() <= 1e-10 * np.abs(x).max()


@pytest.mark.parametrize("nranks", [2, 3, 4, 8])
def test_tree_parallel_ranks_emulated(nranks):
    """The multi-rank path (proportional mapping, staged factorisation with packed boundary fronts, staged
    solve) with every rank emulated as its own fkeep on this GPU: must reproduce the single-GPU result."""
    from galahad_b200.dist_emulate import EmulatedRanks
    for prob in (P.laplace_problem(12), P.kkt_grid_problem(10, dlo=1e-10), P.random_sym_problem(400, 0.01, "indef", seed=5)):
        xs, b = rhs_for(prob)
        ref, flag = gpu_factor(prob)
        f, xr = ref.solve(b)
        er = EmulatedRanks(nranks)
        er.analyse(prob["n"], prob["ptr"], prob["row"], prob["order"])
        assert er.factor(prob["val"]) >= 0
        assert er.num_neg == ref.inform.num_neg
        x = er.solve(b)
        assert P.backward_error(prob, x, b) <= max(1e-14, 10 * P.backward_error(prob, xr, b))
        assert np.abs(x - xr).max() <= 1e-10 * max(1.0, np.abs(xr).max())
        er.free()


def test_solve_block_boundaries_and_partial_jobs():
    """The solve kernels at their seams: fronts of exactly / just below / just above one and several
    128-column blocks (fused single-block launch, step kernels, chain kernels with a partial last block),
    all jobs of datatypes.f90:62-66 on a multi-block front, and bit-reproducibility of repeated solves
    (the chain kernels' CTAs arrive in any order, the summation orders are fixed)."""
    for n in (31, 32, 33, 127, 128, 129, 255, 256, 257, 384, 700):
        A, prob = P.dense_front_problem(n, "spd" if n % 2 else "indef", seed=1000 + n)
        s, flag = gpu_factor(prob)
        assert flag >= 0 and s.inform.num_neg == prob["inertia"][1], n
        b = A @ np.linspace(-1.0, 1.0, n)
        f, x = s.solve(b)
        be = np.abs(A @ x - b).max() / (np.abs(A).sum(1).max() * np.abs(x).max() + np.abs(b).max())
        assert f == 0 and be <= 1e-13, (n, be)
        f, x2 = s.solve(b)
        assert np.array_equal(x, x2), n
        f, y = s.solve(b, job=1)
        f, z = s.solve(y, job=4)
        assert np.abs(z - x).max() <= 1e-11 * max(1, np.abs(x).max()), n
        f, y2 = s.solve(y, job=2)
        f, z2 = s.solve(y2, job=3)
        assert np.abs(z2 - x).max() <= 1e-11 * max(1, np.abs(x).max()), n
        # several right-hand sides are solved column by column: same bits as one at a time
        B = np.stack([b, 2 * b - 1.0, np.ones(n)], axis=1)
        f, X = s.solve(B)
        assert np.array_equal(X[:, 0], x), n
        f, x3 = s.solve(np.ones(n))
        assert np.array_equal(X[:, 2], x3), n


@pytest.mark.parametrize("env", [{"B200LDLT_SOLVE_INV": "0"}, {"B200LDLT_SOLVE_INV": "1"}, {"B200LDLT_SOLVE_PDL": "0"},
                                 {"B200LDLT_SOLVE_T64": "0"}, {"B200LDLT_SOLVE_INV_MB": "0"}, {"B200LDLT_CHAIN": "0"}],
                         ids=lambda e: ",".join(f"{k[9:]}={v}" for k, v in e.items()))
def test_solve_paths_agree(monkeypatch, env):
    """The switchable paths of the solve (substitution / sub-triangle inverses / block inverses in the
    chain kernels, no room for the block inverses, no programmatic dependent launches, 256-thread CTAs for tiny fronts, step
    kernels instead of chains) against the default path: same problem, same right-hand side, solutions equal to rounding
    and each backward stable.  A grid problem (small fronts, fused levels, chains with tails) and a dense indefinite front
    with failed pivots (blocks whose inverses exceed their bounds)."""
    probs = [("grid", None, P.kkt_grid_problem(14)), ("lap", None, P.laplace_problem(24))]
    A, dp = P.dense_front_problem(1300, "indef", seed=7)
    probs.append(("dense", A, dp))
    ref = {}
    for name, A, prob in probs:
        s, flag = gpu_factor(prob)
        assert flag >= 0
        b = np.cos(np.arange(prob["n"]))
        f, x = s.solve(b)
        assert f == 0
        ref[name] = (b, x)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for name, A, prob in probs:
        s, flag = gpu_factor(prob)
        assert flag >= 0
        b, x0 = ref[name]
        f, x = s.solve(b)
        assert f == 0
        if A is None:
            assert P.backward_error(prob, x, b) <= 1e-14, name
        else:
            assert np.abs(A @ x - b).max() / (np.abs(A).sum(1).max() * np.abs(x).max() + np.abs(b).max()) <= 1e-13
        assert np.abs(x - x0).max() <= 1e-9 * np.abs(x0).max(), (name, np.abs(x - x0).max())
        f, x2 = s.solve(b)
        assert np.array_equal(x, x2), name


def test_c2_full_size_properties():
    """BASELINE configs[1] at full size (64^3 7-point Laplacian, n = 262 144): size-independent properties --
    SPD so no negative pivots and no delays, the posdef and the pivoted path agree, backward error <= 1e-14,
    solves reproducible to the bit."""
    prob = P.laplace_problem(64)
    s, flag = gpu_factor(prob)
    assert flag == 0 and s.inform.num_neg == 0 and s.inform.num_delay == 0
    xs, b = rhs_for(prob)
    f, x = s.solve(b)
    assert f == 0 and P.backward_error(prob, x, b) <= 1e-14
    f, x2 = s.solve(b)
    assert np.array_equal(x, x2)
    s2, flag2 = gpu_factor(prob, posdef=True)
    f, xp = s2.solve(b)
    assert flag2 == 0 and P.backward_error(prob, xp, b) <= 1e-14
    assert np.abs(xp - x).max() <= 1e-10 * np.abs(x).max()


def test_value_gather_on_device_matches_host_gather():
    """SLS's value gather (sls.f90:4106-4147) done on the device through the MAPS of
    SLS_coord_to_sorted_csr: entries in arbitrary order, duplicates (up to four per position, summed in input
    order), out-of-range entries (map 0) and positions without an entry (zero).  Must give the bits of the
    host-side gather followed by b200ldlt_factor."""
    rng = np.random.default_rng(77)
    prob = P.kkt_grid_problem(8)
    n, ptr, row, val = prob["n"], prob["ptr"], prob["row"], prob["val"]
    nz = val.size
    # user entries: every position once (shuffled), some split into 2-4 duplicates, a few out-of-range ones
    pos = rng.permutation(nz)
    pos = pos[rng.random(nz) > 0.02]                              # 2 % of the positions get no entry at all
    dup = np.concatenate([rng.choice(pos, nz // 10), rng.choice(pos, nz // 30), rng.choice(pos[:50], 100)])
    entries = np.concatenate([pos, dup])
    rng.shuffle(entries)
    uval = rng.uniform(-1, 1, entries.size)
    # the diagonal stays dominant enough for a stable factorisation
    diag = ptr[:-1]
    seen = np.zeros(nz, bool)
    maps = np.zeros(entries.size + 7, np.int32)
    for l, p in enumerate(entries):
        maps[l] = (p + 1) if not seen[p] else -(p + 1)
        seen[p] = True
    user = np.concatenate([uval, rng.uniform(-1, 1, 7)])          # the last seven are out of range: map 0
    is_diag = np.isin(entries, diag)
    user[:entries.size][is_diag] = np.abs(user[:entries.size][is_diag]) + 8.0
    host = np.zeros(nz)
    for l in range(entries.size):                                  # the reference's loop
        k = maps[l]
        if k > 0:
            host[k - 1] = user[l]
        elif k < 0:
            host[-k - 1] += user[l]
    s1 = Solver()
    assert s1.analyse(n, ptr, row, prob["order"]) >= 0
    assert s1.factor(host) >= 0
    s2 = Solver()
    assert s2.analyse(n, ptr, row, prob["order"]) >= 0
    assert s2.set_value_map(maps) == 0
    assert s2.factor_mapped(user) >= 0
    assert (s1.inform.num_neg, s1.inform.num_two, s1.inform.num_delay) == (s2.inform.num_neg, s2.inform.num_two, s2.inform.num_delay)
    b = rng.uniform(-1, 1, n)
    f1, x1 = s1.solve(b)
    f2, x2 = s2.solve(b)
    assert f1 == 0 and f2 == 0 and np.array_equal(x1, x2)
    # a second factorisation with new values reuses the uploaded map
    user2 = user.copy()
    user2[:entries.size][is_diag] += 1.0
    assert s2.factor_mapped(user2) >= 0
    f3, x3 = s2.solve(b)
    assert f3 == 0 and not np.array_equal(x3, x2)
    # b200ldlt_factor_changed: only the changed entries travel (the interior-point loop); the bits of a full factor_mapped
    which = np.nonzero(np.concatenate([is_diag, np.zeros(7, bool)]))[0].astype(np.int32)
    user3 = user2.copy()
    user3[which] += 0.5
    assert s2.factor_changed(which, user3[which]) >= 0
    f4, x4 = s2.solve(b)
    s3 = Solver()
    assert s3.analyse(n, ptr, row, prob["order"]) >= 0
    assert s3.set_value_map(maps) == 0
    assert s3.factor_changed(which, user3[which]) == -1          # no complete factor_mapped on this handle yet
    assert s3.factor_mapped(user3) >= 0
    f5, x5 = s3.solve(b)
    assert f4 == 0 and f5 == 0 and np.array_equal(x4, x5)
    assert s2.factor_changed(np.array([user.size + 3], np.int32), np.array([1.0])) == -1     # position out of range
    assert s2.factor_changed(np.zeros(0, np.int32), np.zeros(0)) >= 0                        # nothing changed: same factors
    f6, x6 = s2.solve(b)
    assert np.array_equal(x6, x4)


def test_scaled_factorisation():
    """ssids_factor's SCALE argument and options%scaling = 4 (ssids.f90:900-1030, fkeep.F90:247-310): a
    symmetrically mis-scaled KKT matrix factorised with user-supplied factors that undo the damage gives the
    pivot sequence statistics of the well-scaled matrix and solves the mis-scaled system; equilibration gives
    a clean factorisation too, as do the matching-based scalings 1 (Hungarian) and 2 (auction); partial solves compose;
    scaling 3 (saved by a matching-based order) after a plain analyse is SSIDS_ERROR_NO_SAVED_SCALING."""
    rng = np.random.default_rng(21)
    prob = P.kkt_grid_problem(8)
    n, ptr, row = prob["n"], prob["ptr"], prob["row"]
    col = np.repeat(np.arange(n), np.diff(ptr))
    d = 10.0 ** rng.uniform(-5, 5, n)
    bad = dict(prob, val=prob["val"] * d[row] * d[col])                 # D A D
    A = P.full_from_lower(n, ptr, row, bad["val"])
    xs = rng.uniform(-1, 1, n)
    b = A @ xs
    good, flag = gpu_factor(prob)
    # user-supplied scaling 1/d: the engine factorises exactly the well-scaled matrix again (up to rounding of d * (1/d))
    s = Solver()
    assert s.analyse(n, ptr, row, prob["order"]) >= 0
    flag, sc = s.factor_scaled(bad["val"], scale=1.0 / d)
    assert flag >= 0
    assert (s.inform.num_neg, s.inform.num_delay) == (good.inform.num_neg, good.inform.num_delay)
    f, x = s.solve(b)
    assert f == 0 and P.backward_error(bad, x, b) <= 1e-14
    f, y = s.solve(b, job=1)
    f, z = s.solve(y, job=4)
    assert np.abs(z - x).max() <= 1e-12 * np.abs(x).max()
    # equilibration computed by the library, returned to the caller
    flag, sc = s.factor_scaled(bad["val"], scaling=4)
    assert flag >= 0 and sc is not None and sc.min() > 0 and s.inform.num_neg == prob["inertia"][1]
    f, x4 = s.solve(b)
    assert f == 0 and P.backward_error(bad, x4, b) <= 1e-14
    # matching-based scaling by the Hungarian algorithm (options%scaling = 1): |s_i a_ij s_j| <= 1, clean factorisation
    flag, sc1 = s.factor_scaled(bad["val"], scaling=1)
    assert flag >= 0 and sc1 is not None and sc1.min() > 0 and s.inform.num_neg == prob["inertia"][1]
    assert (np.abs(bad["val"]) * sc1[row] * sc1[col]).max() <= 1.0 + 1e-12
    f, x1 = s.solve(b)
    assert f == 0 and P.backward_error(bad, x1, b) <= 1e-14
    # auction algorithm (options%scaling = 2): an approximate matching, entries of the scaled matrix bounded by e
    flag, sc2 = s.factor_scaled(bad["val"], scaling=2)
    assert flag >= 0 and sc2 is not None and sc2.min() > 0 and s.inform.num_neg == prob["inertia"][1]
    assert (np.abs(bad["val"]) * sc2[row] * sc2[col]).max() <= np.e * (1 + 1e-12)
    f, x2 = s.solve(b)
    assert f == 0 and P.backward_error(bad, x2, b) <= 1e-14
    flag, _ = s.factor_scaled(bad["val"], scaling=3)
    assert flag == -15
    # clearing the scaling restores the plain path
    flag, _ = s.factor_scaled(prob["val"], scaling=0)
    b0 = P.full_from_lower(n, ptr, row, prob["val"]) @ xs
    f, x0 = s.solve(b0)
    fg, xg = good.solve(b0)
    assert flag >= 0 and np.array_equal(x0, xg)


def test_matching_based_order_and_its_scaling():
    """ssids_analyse with options%ordering = 2 and ssids_factor with options%scaling = 3 (ssids.f90:321-358, 990-998;
    match_order.f90): a saddle-point matrix with a ZERO (2,2) block and tiny, badly scaled (1,1) diagonals.  Under a
    fill-reducing order that knows nothing of the values many pivots fail their threshold test and are delayed; with
    the matching-based order every constraint sits next to a variable it is matched with (a 2 x 2 pivot at hand) and the
    matching's scaling makes those entries 1: far fewer delays, exact inertia (n_v, m, 0) and a solved system."""
    import scipy.sparse as sp
    rng = np.random.default_rng(17)
    g = 10
    nv = g ** 3
    m = nv // 2
    lap = P.laplace_problem(g)
    Hl = sp.csc_matrix((lap["val"], lap["row"], lap["ptr"]), shape=(nv, nv))
    H = (Hl + sp.tril(Hl, -1).T) * 1e-3 + sp.diags(10.0 ** rng.uniform(-8, -2, nv))
    rows = np.repeat(np.arange(m), 2)
    cols = np.stack([np.arange(m), np.arange(m) + m], 1).ravel()
    J = sp.csr_matrix((10.0 ** rng.uniform(-2, 2, 2 * m), (rows, cols)), shape=(m, nv))
    K = sp.bmat([[H, J.T], [J, None]], format="csc")
    n, ptr, row, val = P.lower_csc(K)
    prob = dict(n=n, ptr=ptr, row=row, val=val)
    xs = rng.uniform(-1, 1, n)
    b = K @ xs
    # values-blind order: nested dissection of the pattern
    L = _lib.lib()
    plain = Solver()
    o = np.zeros(n, np.int32)
    assert L.b200ldlt_order(n, _lib.as_ptr(ptr.astype(np.int64), C.c_int64), _lib.as_ptr(row, C.c_int), 3, _lib.as_ptr(o, C.c_int),
                            C.byref(plain.opt)) == 0
    assert plain.analyse(n, ptr, row, o) >= 0
    assert plain.factor(val) >= 0
    d_plain = plain.inform.num_delay
    assert (plain.inform.num_neg, plain.inform.matrix_rank) == (m, n)
    # matching-based order + its scaling
    s = Solver()
    assert s.analyse_matched(n, ptr, row, val) == 0
    assert sorted(s.order) == list(range(n))
    flag, sc = s.factor_scaled(val, scaling=3)
    assert flag >= 0 and sc is not None and sc.min() > 0
    col = np.repeat(np.arange(n), np.diff(ptr))
    assert (np.abs(val) * sc[row] * sc[col]).max() <= 1 + 1e-9
    assert (s.inform.num_neg, s.inform.matrix_rank) == (m, n)
    assert s.inform.num_delay * 4 <= max(d_plain, 4), (s.inform.num_delay, d_plain)
    f, x = s.solve(b)
    # backward stable for the matrix that was factorised, S K S (the reference engine: 3e-16 there, 6e-13 on K itself)
    assert f == 0 and P.backward_error(dict(prob, val=val * sc[row] * sc[col]), x / sc, sc * b) <= 1e-14
    assert P.backward_error(prob, x, b) <= 1e-11
    # the saved scaling belongs to the analyse: a new plain analyse forgets it
    assert s.analyse(n, ptr, row, o) >= 0
    flag, _ = s.factor_scaled(val, scaling=3)
    assert flag == -15


def _arms_problem(arms, ns, kind, seed=3):
    """Dense arms of different widths joined by a dense separator: one level of fronts with m > n and different
    block counts, then the separator front."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    n = sum(arms) + ns
    A = np.zeros((n, n))

    def sym(k):
        M = rng.uniform(-1, 1, (k, k))
        return (M + M.T) / 2
    o = 0
    for a in arms:
        A[o:o + a, o:o + a] = sym(a)
        A[n - ns:, o:o + a] = rng.uniform(-1, 1, (ns, a))
        o += a
    A[n - ns:, n - ns:] = sym(ns)
    A = np.tril(A)
    A = A + np.tril(A, -1).T
    if kind == "spd":
        A += n * np.eye(n)
    nn, ptr, row, val = P.lower_csc(sp.csc_matrix(A))
    neg = int((np.linalg.eigvalsh(A) < 0).sum())
    return A, dict(n=n, ptr=ptr, row=row, val=val, order=np.arange(n, dtype=np.int32), inertia=(n - neg, neg, 0))


@pytest.mark.parametrize("defer", [4, 2, 3])
def test_deferred_updates(monkeypatch, defer):
    """The look-ahead schedule of the large fronts (k_upd3): the columns behind a group of `defer` blocks get the
    group's pivots in one trailing update of rank defer * 128, cut into a part on the critical path and slices beside
    the next blocks (B200LDLT_DEFER_MINM lowers the front size from which groups are used).  Fronts with a short last
    block, several fronts of different block counts in one launch, failed pivots anywhere in a group (stall + closing
    flush), and a sparse problem; against the eigenvalues / the run without groups."""
    monkeypatch.setenv("B200LDLT_DEFER_MINM", "0")
    monkeypatch.setenv("B200LDLT_DEFER", str(defer))
    for arms, ns, kind in (((300, 200, 520), 900, "spd"), ((300,), 800, "indef"), ((520, 300), 1000, "indef"),
                           ((129, 257), 1000, "spd"), ((640,), 1100, "indef"), ((1100, 700), 1300, "spd")):
        A, prob = _arms_problem(arms, ns, kind)
        s, flag = gpu_factor(prob)
        assert flag >= 0 and s.inform.num_neg == prob["inertia"][1], (arms, kind)
        b = A @ np.ones(prob["n"])
        f, x = s.solve(b)
        be = np.abs(A @ x - b).max() / (np.abs(A).sum(1).max() * np.abs(x).max() + np.abs(b).max())
        assert f == 0 and be <= 1e-13, (arms, kind, be)
    for kind, n in (("indef", 1300), ("zdiag", 2100), ("kkt", 1500)):
        A, prob = P.dense_front_problem(n, kind, seed=n)
        s, flag = gpu_factor(prob)
        assert flag >= 0 and s.inform.num_neg == prob["inertia"][1], (kind, n)
        b = A @ np.ones(n)
        f, x = s.solve(b)
        be = np.abs(A @ x - b).max() / (np.abs(A).sum(1).max() * np.abs(x).max() + np.abs(b).max())
        assert f == 0 and be <= 1e-13, (kind, n, be)
    prob = P.laplace_problem(32)
    xs, b = rhs_for(prob)
    s, flag = gpu_factor(prob)
    f, x = s.solve(b)
    assert flag == 0 and P.backward_error(prob, x, b) <= 1e-14
    monkeypatch.setenv("B200LDLT_DEFER", "1")
    s0, flag0 = gpu_factor(prob)
    f, x0 = s0.solve(b)
    assert np.abs(x - x0).max() <= 1e-12 * np.abs(x0).max()
    monkeypatch.setenv("B200LDLT_UPD3", "0")       # the small-front kernel on every front
    s1, flag1 = gpu_factor(prob)
    f, x1 = s1.solve(b)
    assert flag1 == 0 and np.abs(x - x1).max() <= 1e-12 * np.abs(x1).max()


def test_reanalyse_on_a_live_handle():
    """analyse, factor, analyse, factor on one Solver: the fkeep made from the first akeep outlives it until the next
    factorisation replaces it (the akeep is shared with its fkeeps); freeing akeep before fkeep is legal as well."""
    p1, p2 = P.laplace_problem(6), P.kkt_grid_problem(6)
    s = Solver()
    for prob in (p1, p2, p1):
        assert s.analyse(prob["n"], prob["ptr"], prob["row"], prob["order"]) >= 0
        assert s.factor(prob["val"]) >= 0
        xs, b = rhs_for(prob)
        f, x = s.solve(b)
        assert f >= 0 and P.backward_error(prob, x, b) <= 1e-14 and s.inform.num_neg == prob["inertia"][1]
    import ctypes as C
    none = C.c_void_p()
    s.L.b200ldlt_free(C.byref(s.akeep), C.byref(none))       # akeep first ...
    assert not s.akeep.value
    s.L.b200ldlt_free_fkeep(C.byref(s.fkeep))                # ... its fkeep afterwards
    assert not s.fkeep.value


def _nodes(prob, posdef=False):
    s, flag = gpu_factor(prob, posdef)
    nn = s.numeric_tree()["nnodes"]
    return s, flag, [s.node(i) for i in range(nn)]


@pytest.mark.parametrize("case", ["lap7_16", "kkt7_12_hard", "qpband_20000", "rand_indef", "dense_kkt_64", "dense_zdiag_40"])
def test_small_fronts_in_one_launch(monkeypatch, case):
    """k_front (fronts of at most 64 candidate columns assembled, factorised and their contribution block formed by one
    CTA, front_kernels.cuh) against the general path (B200LDLT_FRONT=0) on the same matrix: same inertia, delays and
    pivot counts; fronts of at most 32 columns whose subtree has no wider front are equal to the bit (same operations in
    the same order); the rest agree to rounding; fronts with a failed pivot are rolled back and refactorised by the
    general path's retry rounds (rand_indef, dense_*: delays and 2x2 pivots)."""
    prob = {"lap7_16": lambda: P.laplace_problem(16), "kkt7_12_hard": lambda: P.kkt_grid_problem(12, dlo=1e-10),
            "qpband_20000": lambda: P.qpband_problem(20000), "rand_indef": lambda: P.random_sym_problem(2000, 0.004, "indef", seed=5),
            "dense_kkt_64": lambda: P.dense_front_problem(64, "kkt")[1], "dense_zdiag_40": lambda: P.dense_front_problem(40, "zdiag")[1]}[case]()
    monkeypatch.setenv("B200LDLT_FRONT", "0")
    s0, f0, n0 = _nodes(prob)
    monkeypatch.setenv("B200LDLT_FRONT", "1")
    s1, f1, n1 = _nodes(prob)
    assert f0 == f1 and f1 >= 0
    assert (s0.inform.num_neg, s0.inform.num_delay, s0.inform.num_two) == (s1.inform.num_neg, s1.inform.num_delay, s1.inform.num_two)
    if prob.get("inertia") is not None:
        assert s1.inform.num_neg == prob["inertia"][1]
    assert s1.inform.reserved[0] < s0.inform.reserved[0] or case.startswith("dense") or case == "rand_indef"   # fewer launches
    nbit = 0
    for a, b in zip(n0, n1):
        assert (a["m"], a["n"], a["nelim"]) == (b["m"], b["n"], b["nelim"])
        assert np.array_equal(a["perm"], b["perm"])
        if np.array_equal(a["L"], b["L"]) and np.array_equal(a["d"], b["d"]):
            nbit += 1
        else:
            scale = max(1.0, float(np.abs(a["L"]).max()))
            assert np.abs(a["L"] - b["L"]).max() <= 1e-9 * scale
    if case == "lap7_16":
        assert nbit >= len(n0) // 2          # the chains of 32-column fronts at the bottom of the tree
    xs, b = rhs_for(prob)
    f, x0 = s0.solve(b)
    f, x1 = s1.solve(b)
    assert P.backward_error(prob, x1, b) <= max(1e-14, 10 * P.backward_error(prob, x0, b))


def test_c4_qpband_full_size_properties():
    """BASELINE configs[3] at full size (QPBAND, N = 10^6: 45 330 fronts of at most 49 rows, all through k_front):
    inertia (N, N/2, 0) exact, no delays, backward error, and a refactorisation with new barrier diagonals on the same
    handle (the interior-point loop's pattern)."""
    prob = P.qpband_problem(1000000)
    s, flag = gpu_factor(prob)
    assert flag == 0 and s.inform.num_neg == prob["inertia"][1] and s.inform.num_delay == 0
    xs, b = rhs_for(prob)
    f, x = s.solve(b)
    assert f == 0 and P.backward_error(prob, x, b) <= 1e-14
    prob2 = P.qpband_problem(1000000, it=7)
    assert s.factor(prob2["val"]) == 0 and s.inform.num_neg == prob2["inertia"][1]
    xs, b2 = rhs_for(prob2)
    f, x2 = s.solve(b2)
    assert f == 0 and P.backward_error(prob2, x2, b2) <= 1e-14
