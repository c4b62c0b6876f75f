"""C5 on one GPU: backward error before and after one step of iterative refinement (residual in fp64 on the host)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from galahad_b200 import problems as P
from galahad_b200._lib import Solver
g = int(sys.argv[1]) if len(sys.argv) > 1 else 160
prob = P.laplace_problem(g, stencil=27)
s = Solver(); s.analyse(prob["n"], prob["ptr"], prob["row"], prob["order"])
assert s.factor(prob["val"]) >= 0
A = P.full_from_lower(prob["n"], prob["ptr"], prob["row"], prob["val"])
rng = np.random.default_rng(1); xs = rng.uniform(-1, 1, prob["n"]); b = A @ xs
f, x = s.solve(b)
def berr(x):
    r = A @ x - b
    return float(np.abs(r).max() / (abs(A).sum(axis=1).max() * np.abs(x).max() + np.abs(b).max()))
print("berr unrefined", berr(x))
f, dx = s.solve(b - A @ x)
x1 = x + dx
print("berr after one refinement", berr(x1))
