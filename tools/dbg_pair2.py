"""Two dense arms joined by a dense separator: fronts with m > n and several blocks (look-ahead + pair updates)."""
import os, sys
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from galahad_b200 import problems as P
from galahad_b200._lib import Solver
n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 300
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 800
kind = sys.argv[3] if len(sys.argv) > 3 else "spd"
rng = np.random.default_rng(3)
n = 2 * n1 + ns
A = np.zeros((n, n))
def sym(k): 
    M = rng.uniform(-1, 1, (k, k)); return (M + M.T) / 2
A[:n1, :n1] = sym(n1); A[n1:2*n1, n1:2*n1] = sym(n1); A[2*n1:, 2*n1:] = sym(ns)
A[2*n1:, :n1] = rng.uniform(-1, 1, (ns, n1)); A[2*n1:, n1:2*n1] = rng.uniform(-1, 1, (ns, n1))
A = np.tril(A); A = A + np.tril(A, -1).T
if kind == "spd":
    A += n * np.eye(n)
nn, ptr, row, val = P.lower_csc(sp.csc_matrix(A))
prob = dict(n=n, ptr=ptr, row=row, val=val, order=np.arange(n, dtype=np.int32))
s = Solver(device=0)
s.analyse(n, ptr, row, prob["order"])
nt = s.numeric_tree()
print("nodes", nt["nnodes"], "sptr", nt["sptr"][:6], "rows", np.diff(nt["rptr"])[:6])
f = s.factor(val)
b = A @ np.ones(n)
fl, x = s.solve(b)
neg = int((np.linalg.eigvalsh(A) < 0).sum())
print(f"PAIR={os.environ.get('B200LDLT_PAIR','1')} n1={n1} ns={ns} {kind}: flag {f} neg {s.inform.num_neg} (expect {neg}) delay {s.inform.num_delay} "
      f"berr {np.abs(A@x-b).max()/(np.abs(A).sum(1).max()*np.abs(x).max()+np.abs(b).max()):.2e}")
