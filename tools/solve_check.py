"""Diagnostic: backward error and solve time on the dense-front fixtures (and, with --big, C2 and a 40^3 KKT grid)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from galahad_b200 import problems as P
from galahad_b200._lib import Solver

def run(prob, A=None, posdef=False):
    s = Solver(device=0)
    s.analyse(prob["n"], prob["ptr"], prob["row"], prob["order"])
    flag = s.factor(prob["val"], posdef=posdef)
    n = prob["n"]
    if A is None:
        A = P.full_from_lower(n, prob["ptr"], prob["row"], prob["val"])
    b = A @ np.ones(n)
    ts = []
    for _ in range(3):
        f, x = s.solve(b)
        ts.append(s.inform.time_solve * 1e3)
    r = A @ x - b
    be = np.abs(r).max() / (abs(A).sum(1).max() * np.abs(x).max() + np.abs(b).max())
    print(f"{prob.get('name','?')} n={n} flag={flag} neg={s.inform.num_neg} two={s.inform.num_two} "
          f"delay={s.inform.num_delay} berr={be:.2e} factor={s.inform.time_factor*1e3:.2f}ms solve_ms={min(ts):.3f}", flush=True)
    s.free()

for kind, n in (("zdiag", 142), ("kkt", 517), ("indef", 1300), ("zdiag", 2100), ("spd", 1100)):
    A, prob = P.dense_front_problem(n, kind, seed=n)
    prob["name"] = f"dense_{kind}"
    run(prob, A)
if "--big" in sys.argv:
    prob = P.laplace_problem(64); prob["name"] = "lap64"
    run(prob, A=P.full_from_lower(prob["n"], prob["ptr"], prob["row"], prob["val"]))
    prob = P.kkt_grid_problem(40); prob["name"] = "kkt40"
    run(prob, A=P.full_from_lower(prob["n"], prob["ptr"], prob["row"], prob["val"]))
