"""Development check: dense symmetric matrices factorised as ONE front (identity order), several
orders and kinds; prints inertia vs numpy and the scaled backward error."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from galahad_b200._lib import Solver

def dense_problem(n, kind, seed):
    rng = np.random.default_rng(seed)
    A = rng.uniform(-1, 1, (n, n)); A = (A + A.T) / 2
    if kind == "zdiag":
        np.fill_diagonal(A, 0.0)
    elif kind == "kkt":
        k = n // 3
        A[n - k:, n - k:] = 0.0
        A[:n - k, :n - k] += np.eye(n - k) * 2
    elif kind == "spd":
        A = A @ A.T + n * np.eye(n)
    elif kind == "small":
        A[:, : n // 2] *= 1e-3; A[: n // 2, :] *= 1e-3
    ptr = [0]; row = []; val = []
    for j in range(n):
        for i in range(j, n):
            row.append(i); val.append(A[i, j])
        ptr.append(len(row))
    return A, np.array(ptr, np.int64), np.array(row, np.int32), np.array(val)

ok = True
kinds = sys.argv[1].split(",") if len(sys.argv) > 1 else ("spd", "indef", "zdiag", "kkt", "small")
sizes = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else (20, 33, 64, 65, 128, 129, 200, 300, 517)
for kind in kinds:
    for n in sizes:
        A, ptr, row, val = dense_problem(n, kind, n)
        s = Solver()
        s.analyse(n, ptr, row, np.arange(n, dtype=np.int32))
        f = s.factor(val)
        inf = s.inform.asdict()
        b = A @ np.ones(n)
        fs, x = s.solve(b)
        be = np.abs(A @ x - b).max() / (np.abs(A).sum(1).max() * np.abs(x).max() + np.abs(b).max())
        neg = int((np.linalg.eigvalsh(A) < 0).sum())
        good = f >= 0 and be < 1e-13 and neg == inf["num_neg"]
        ok &= good
        print(f"{kind:6s} n={n:4d} flag={f} neg={inf['num_neg']}/{neg} two={inf['num_two']} delay={inf['num_delay']} "
              f"nfp={inf['not_first_pass']} nsp={inf['not_second_pass']} berr={be:.2e} {'' if good else '  <-- BAD'}", flush=True)
        s.free()
print("ALL OK" if ok else "SOME FAILED")

def residual_map(kind, n):
    """Which entries of P A P^T - L D L^T are wrong (single front)."""
    A, ptr, row, val = dense_problem(n, kind, n)
    s = Solver()
    s.analyse(n, ptr, row, np.arange(n, dtype=np.int32))
    s.factor(val)
    g = s.node(0)
    L, d, perm, ne = g["L"], g["d"], g["perm"], g["nelim"]
    print("m n nelim", g["m"], g["n"], ne, "inform", {k: s.inform.asdict()[k] for k in ("num_two", "not_first_pass", "num_delay")})
    Dm = np.zeros((ne, ne))
    i = 0
    while i < ne:
        if i + 1 < ne and np.isinf(d[2 * i + 2]):
            Dinv = np.array([[d[2 * i], d[2 * i + 1]], [d[2 * i + 1], d[2 * i + 3]]])
            Dm[i:i + 2, i:i + 2] = np.linalg.inv(Dinv)
            i += 2
        else:
            Dm[i, i] = 1.0 / d[2 * i] if d[2 * i] != 0 else 0.0
            i += 1
    Lf = np.tril(L[:, :ne], -1)[:ne] + np.eye(ne)
    R = A[np.ix_(perm[:ne], perm[:ne])] - Lf @ Dm @ Lf.T
    bad = np.abs(R) > 1e-10
    print("bad entries", int(bad.sum()))
    rows = np.where(bad.any(1))[0]; cols = np.where(bad.any(0))[0]
    print("bad rows", rows[:40], "\nbad cols", cols[:40])
    print("perm sorted ok", sorted(perm[:ne].tolist()) == list(range(ne)))
    np.set_printoptions(linewidth=250, precision=1)
    edges = sorted(set([0] + [int(v) for v in sys.argv[4].split(",")] + [ne])) if len(sys.argv) > 4 else list(range(0, ne, 32)) + [ne]
    print("edges", edges)
    G = np.array([[np.abs(R[edges[a]:edges[a + 1], edges[b]:edges[b + 1]]).max() for b in range(len(edges) - 1)] for a in range(len(edges) - 1)])
    print(G)
    print("two-by-two firsts", [i for i in range(ne - 1) if np.isinf(d[2 * i + 2])][:80])

if len(sys.argv) > 3:
    residual_map(sys.argv[1].split(",")[0], int(sys.argv[2].split(",")[0]))
