"""The SLS host interface for solver 'b200ldlt', exercised the way the reference exercises its own:
src/sls/C/slst.c (three storage schemes, 1 RHS, iterative refinement, partial solves) and
src/sls/slst.f90:29-56 (5x5 KATs, provided reverse order)."""
import numpy as np
import pytest

from galahad_b200 import problems as P
from galahad_b200.sls import SLS, coord_to_sorted_csr
from util import golden

EPS = np.finfo(float).eps
K = golden("kat_sls.json")
N = 5
ROW = np.array(K["indef"]["row"]) - 1            # slst.c:23-30 holds the same data 0-based
COL = np.array(K["indef"]["col"]) - 1
PTR = np.array([0, 1, 3, 5, 6, 7])
VAL = np.array(K["indef"]["val"], float)
DENSE = np.array([2.0, 3.0, 0.0, 0.0, 4.0, 1.0, 0.0, 0.0, 5.0, 0.0, 0.0, 6.0, 0.0, 0.0, 1.0])
RHS = np.array(K["indef"]["rhs"], float)
SOL = np.array(K["indef"]["sol"], float)


def test_unknown_solver_and_defaults():
    s = SLS("sils")
    assert s.status.value == -26                                   # GALAHAD_error_unknown_solver
    s = SLS("B200LDLT")
    c = s.control
    assert s.status.value == 0
    # sls.f90:129-357 defaults (+ ordering -1 for the ssids class, :887-892)
    assert (c.pivot_control, c.node_amalgamation, c.relative_pivot_tolerance, c.max_iterative_refinements) == (1, 32, 0.01, 0)
    assert c.absolute_pivot_tolerance == EPS and c.ordering == -1


def test_coord_to_sorted_csr_counts_and_map():
    # duplicates, an upper-triangle entry, an out-of-range entry; row 4 has no entry at all, which
    # the reference does not count as a missing diagonal (the test sits in the size > 0 branch,
    # sls.f90:8529-8547) although it still reserves the diagonal slot
    row = np.array([1, 2, 2, 3, 1, 2, 9, 4, 3])
    col = np.array([1, 1, 2, 3, 2, 1, 1, 3, 3])
    m, p, cnt = coord_to_sorted_csr(4, row, col)
    assert cnt == dict(dup=3, oor=1, upper=1, missing_diagonals=0)
    # a row with off-diagonal entries only IS counted
    m2, p2, cnt2 = coord_to_sorted_csr(3, np.array([2, 3, 3]), np.array([1, 1, 3]))
    assert cnt2["missing_diagonals"] == 1 and p2.tolist() == [1, 4, 5, 6] and m2.tolist() == [2, 3, 5]
    assert p.tolist() == [1, 3, 4, 6, 7]
    # (2,1), (1,2) and the second (2,1) all land on slot 2; the smallest entry index owns it
    assert m.tolist() == [1, 2, 3, 4, -2, -2, 0, 5, -4]
    # every row starts with its diagonal slot
    assert all(p[i] <= p[i + 1] - 1 for i in range(4))


def test_analyse_reports_structure_without_gpu():
    s = SLS()
    s.control.f_indexing = False
    assert s.analyse_matrix(N, "coordinate", 7, ROW, COL) == 0
    inf = s.information()
    assert (inf.entries, inf.duplicates, inf.out_of_range, inf.upper, inf.missing_diagonals) == (7, 0, 0, 1, 1)   # entry (2,5) of slst.f90:29-33 is given in the upper triangle
    assert inf.rank == 5 and inf.nodes_assembly_tree >= 1 and inf.entries_in_factors >= 8
    assert s.analyse_matrix(0, "coordinate", 0, ROW, COL) == -3      # n > 0 restriction (sls.h status -3)
    assert s.analyse_matrix(N, "banded", 0, ROW, COL) == -3


@pytest.mark.gpu
@pytest.mark.parametrize("storage", ["coordinate", "sparse_by_rows", "dense"])
def test_slst_c_flow(storage):
    s = SLS()
    s.control.f_indexing = False
    if storage == "coordinate":
        assert s.analyse_matrix(N, storage, 7, ROW, COL) == 0
        assert s.factorize_matrix(VAL) == 0
    elif storage == "sparse_by_rows":
        assert s.analyse_matrix(N, storage, 7, None, COL, PTR) == 0
        assert s.factorize_matrix(VAL) == 0
    else:
        assert s.analyse_matrix(N, storage, 15) == 0
        assert s.factorize_matrix(DENSE) == 0
    good_x = EPS ** 0.3333                                           # slst.c:36
    st, x = s.solve_system(RHS)
    assert st == 0 and np.abs(x - SOL).max() < good_x
    inf = s.information()
    assert inf.status == 0 and inf.rank == 5 and inf.negative_eigenvalues == 2
    s.control.max_iterative_refinements = 1                          # resolve with refinement
    s.reset_control()
    st, x = s.solve_system(RHS)
    assert st == 0 and np.abs(x - SOL).max() < good_x
    st, y = s.partial_solve_system("L", RHS)                         # by parts
    st, y = s.partial_solve_system("D", y)
    st, y = s.partial_solve_system("U", y)
    assert st == 0 and np.abs(y - SOL).max() < good_x
    s.terminate()


@pytest.mark.gpu
def test_provided_order_and_definite_path():
    d = K["spd"]
    row, col, val = np.array(d["row"]) - 1, np.array(d["col"]) - 1, np.array(d["val"], float)
    perm = np.array(K["order_reverse"]) - 1                          # slst.f90:54-56
    for pc in (1, 2):
        s = SLS()
        s.control.pivot_control = pc
        assert s.analyse_matrix(N, "coordinate", 7, row, col, perm=perm) == 0
        assert s.factorize_matrix(val) == 0
        st, x = s.solve_system(np.array(d["rhs"], float))
        assert np.abs(x - np.array(d["sol"], float)).max() < EPS ** 0.5
        assert s.information().negative_eigenvalues == 0
    s = SLS()
    s.control.pivot_control = 2                                      # indefinite matrix on the definite path
    s.analyse_matrix(N, "coordinate", 7, ROW, COL)
    # SSIDS flag -6 (not positive definite) goes through SLS_copy_inform_from_ssids, which files it under
    # GALAHAD_error_restrictions = -3 (sls.f90:1768-1769); the facade follows the reference's table
    assert s.factorize_matrix(VAL) == -3


@pytest.mark.gpu
def test_duplicates_are_summed():
    # split every value of the KAT into two entries: same solution
    row2, col2 = np.concatenate([ROW, ROW]), np.concatenate([COL, COL])
    val2 = np.concatenate([0.25 * VAL, 0.75 * VAL])
    s = SLS()
    assert s.analyse_matrix(N, "coordinate", 14, row2, col2) == 0
    assert s.information().duplicates == 7
    assert s.factorize_matrix(val2) == 0
    st, x = s.solve_system(RHS)
    assert np.abs(x - SOL).max() < EPS ** 0.5


@pytest.mark.gpu
def test_kkt_through_sls_with_nd_order():
    prob = P.kkt_grid_problem(8)
    n = prob["n"]
    colidx = np.repeat(np.arange(n), np.diff(prob["ptr"]))
    s = SLS()
    # lower triangle in coordinate form, shuffled, with the geometric ND order as PERM
    rng = np.random.default_rng(3)
    q = rng.permutation(prob["row"].size)
    assert s.analyse_matrix(n, "coordinate", q.size, prob["row"][q], colidx[q], perm=prob["order"]) == 0
    assert s.factorize_matrix(prob["val"][q]) == 0
    inf = s.information()
    assert inf.negative_eigenvalues == prob["inertia"][1] and inf.rank == n
    xs = rng.uniform(-1, 1, n)
    b = P.full_from_lower(n, prob["ptr"], prob["row"], prob["val"]) @ xs
    st, x = s.solve_system(b)
    assert P.backward_error(prob, x, b) <= 1e-14


@pytest.mark.gpu
def test_iterative_refinement_on_device():
    """SLS_solve_ir (sls.f90:4751-4947) runs on the device: same iterates as the loop written out with
    unrefined solves and a residual over the caller's entries (here with duplicates and a relaxed pivot
    threshold so that refinement has something to do), iteration count reported, early stop honoured."""
    prob = P.kkt_grid_problem(8, dlo=1e-10)
    n = prob["n"]
    colidx = np.repeat(np.arange(n), np.diff(prob["ptr"]))
    rng = np.random.default_rng(11)
    # split a tenth of the entries into two duplicates each
    row, col, val = prob["row"].copy(), colidx.copy(), prob["val"].copy()
    pick = rng.choice(row.size, row.size // 10, replace=False)
    row = np.concatenate([row, row[pick]]); col = np.concatenate([col, col[pick]])
    half = val[pick] * rng.uniform(0.2, 0.8, pick.size)
    val[pick] -= half
    val = np.concatenate([val, half])
    q = rng.permutation(row.size)
    row, col, val = row[q], col[q], val[q]
    A = P.full_from_lower(n, prob["ptr"], prob["row"], prob["val"])
    b = A @ rng.uniform(-1, 1, n)

    def residual(x):                      # the reference's loop over the entries, in input order
        r = b.copy()
        for l in range(row.size):
            i, j, v = row[l], col[l], val[l]
            r[i] -= v * x[j]
            if i != j:
                r[j] -= v * x[i]
        return r

    s = SLS()
    s.control.f_indexing = False
    s.control.relative_pivot_tolerance = 1e-8
    assert s.analyse_matrix(n, "coordinate", row.size, row, col, perm=prob["order"]) == 0
    assert s.factorize_matrix(val) == 0
    st, x0 = s.solve_system(b)                                     # no refinement
    # two refinements written out with plain solves
    x = x0.copy()
    for _ in range(2):
        st, dx = s.solve_system(residual(x))
        x = x + dx
    s.control.max_iterative_refinements = 2
    s.control.acceptable_residual_relative = 0.0
    s.control.acceptable_residual_absolute = 0.0
    s.reset_control()
    st, xr = s.solve_system(b)
    assert st == 0 and s.information().iterative_refinements == 2
    assert np.abs(xr - x).max() <= 1e-13 * max(1.0, np.abs(x).max())
    assert np.abs(residual(xr)).max() <= np.abs(residual(x0)).max()
    # a generous tolerance stops after the first solve
    s.control.acceptable_residual_relative = 1.0
    s.reset_control()
    st, x1 = s.solve_system(b)
    assert st == 0 and s.information().iterative_refinements == 0
    assert np.array_equal(x1, x0)
    s.terminate()


@pytest.mark.gpu
def test_matching_based_scaling_through_sls():
    """control%scaling = -1 (SSIDS's Hungarian scaling, sls.f90:1405-1413): a symmetrically mis-scaled KKT matrix (scale
    factors over sixteen orders of magnitude) factorises with the exact inertia and solves to backward error 1e-14;
    -2 (auction) likewise; -3 (matching-order) stays GALAHAD_unavailable_option."""
    prob = P.kkt_grid_problem(8)
    n = prob["n"]
    colidx = np.repeat(np.arange(n), np.diff(prob["ptr"]))
    rng = np.random.default_rng(5)
    d = 10.0 ** rng.uniform(-8, 8, n)
    val = prob["val"] * d[prob["row"]] * d[colidx]
    bad = dict(prob, val=val)
    s = SLS()
    s.control.f_indexing = False
    s.control.scaling = -1
    assert s.analyse_matrix(n, "coordinate", val.size, prob["row"], colidx, perm=prob["order"]) == 0
    assert s.factorize_matrix(val) == 0
    inf = s.information()
    assert inf.negative_eigenvalues == prob["inertia"][1] and inf.rank == n
    b = P.full_from_lower(n, prob["ptr"], prob["row"], val) @ rng.uniform(-1, 1, n)
    st, x = s.solve_system(b)
    assert st == 0 and P.backward_error(bad, x, b) <= 1e-14
    for sc, want in ((-2, 0), (-3, -3)):            # auction scaling; matching-order scaling: nothing saved -> error_restrictions
        s2 = SLS()
        s2.control.f_indexing = False
        s2.control.scaling = sc
        assert s2.analyse_matrix(n, "coordinate", val.size, prob["row"], colidx, perm=prob["order"]) == 0
        assert s2.factorize_matrix(val) == want
        if want == 0:
            assert s2.information().negative_eigenvalues == prob["inertia"][1]
            st, x2 = s2.solve_system(b)
            assert st == 0 and P.backward_error(bad, x2, b) <= 1e-14


def _sorted_csr_reference(n, row, col):
    """What SLS_coord_to_sorted_csr (sls.f90:8409-8578) must return, stated from its contract rather than from its loops:
    upper triangle by rows, every row led by its diagonal slot (present or not), then its distinct columns ascending;
    the first entry of a position (in input order) owns it, later ones carry minus the position, out-of-range entries 0."""
    ne = len(row)
    inr = (row >= 1) & (row <= n) & (col >= 1) & (col <= n)
    r = np.minimum(row, col)
    c = np.maximum(row, col)
    cols = [set() for _ in range(n + 1)]
    for l in range(ne):
        if inr[l] and c[l] != r[l]:
            cols[r[l]].add(int(c[l]))
    ptr = np.ones(n + 1, np.int64)
    slot = {}
    k = 1
    for i in range(1, n + 1):
        ptr[i - 1] = k
        slot[(i, i)] = k
        k += 1
        for cc in sorted(cols[i]):
            slot[(i, cc)] = k
            k += 1
    ptr[n] = k
    m = np.zeros(ne, np.int64)
    seen = set()
    for l in range(ne):
        if not inr[l]:
            continue
        key = (int(r[l]), int(c[l]))
        m[l] = slot[key] if key not in seen else -slot[key]
        seen.add(key)
    has_any = {int(v) for v in r[inr]}
    has_diag = {int(v) for v in r[inr & (row == col)]}
    cnt = dict(dup=int((m < 0).sum()), oor=int((~inr).sum()), upper=int((inr & (row < col)).sum()),
               missing_diagonals=len(has_any - has_diag))
    return m, ptr, cnt


def test_coord_to_sorted_csr_random_inputs():
    """300 random coordinate lists (duplicates, both triangles, out-of-range indices, rows without a diagonal or without any
    entry) against the contract above: MAPS, PTR and the four counters exactly."""
    rng = np.random.default_rng(8409)
    for t in range(300):
        n = int(rng.integers(1, 30))
        ne = int(rng.integers(0, 4 * n + 2))
        row = rng.integers(1, n + 1, ne)
        col = rng.integers(1, n + 1, ne)
        if t % 3 == 0 and ne > 2:                      # out of range on either side
            bad = rng.choice(ne, size=max(1, ne // 6), replace=False)
            row[bad[::2]] = rng.choice([0, -3, n + 1, n + 7], size=len(bad[::2]))
            col[bad[1::2]] = rng.choice([0, n + 1], size=len(bad[1::2]))
        if t % 4 == 0 and ne > 0:                      # many repeats of a few positions
            k = rng.integers(0, ne, ne // 2)
            row[: len(k)] = row[k]
            col[: len(k)] = col[k]
        m, p, cnt = coord_to_sorted_csr(n, row, col)
        m0, p0, cnt0 = _sorted_csr_reference(n, row, col)
        assert cnt == cnt0, (t, cnt, cnt0)
        assert np.array_equal(np.asarray(p, np.int64), p0), t
        assert np.array_equal(np.asarray(m, np.int64), m0), t
