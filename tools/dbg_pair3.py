"""Several dense arms of different widths joined by a dense separator: one level with fronts of different block counts."""
import os, sys
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from galahad_b200 import problems as P
from galahad_b200._lib import Solver
arms = [int(a) for a in sys.argv[1].split(",")]
ns = int(sys.argv[2])
rng = np.random.default_rng(3)
n = sum(arms) + ns
A = np.zeros((n, n))
def sym(k):
    M = rng.uniform(-1, 1, (k, k)); return (M + M.T) / 2
o = 0
for a in arms:
    A[o:o+a, o:o+a] = sym(a)
    A[n-ns:, o:o+a] = rng.uniform(-1, 1, (ns, a))
    o += a
A[n-ns:, n-ns:] = sym(ns)
A = np.tril(A); A = A + np.tril(A, -1).T + n * np.eye(n)
nn, ptr, row, val = P.lower_csc(sp.csc_matrix(A))
s = Solver(device=0)
s.analyse(n, ptr, row, np.arange(n, dtype=np.int32))
nt = s.numeric_tree()
f = s.factor(val)
b = A @ np.ones(n)
fl, x = s.solve(b)
print(f"PAIR={os.environ.get('B200LDLT_PAIR','1')} arms={arms} ns={ns}: nodes {nt['nnodes']} cols {np.diff(nt['sptr']).tolist()} rows {np.diff(nt['rptr']).tolist()} flag {f} "
      f"berr {np.abs(A@x-b).max()/(np.abs(A).sum(1).max()*np.abs(x).max()+np.abs(b).max()):.2e}")
