import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from galahad_b200 import problems as P
from galahad_b200._lib import Solver
from galahad_b200.dist_emulate import EmulatedRanks
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 12
nr = int(sys.argv[2]) if len(sys.argv) > 2 else 2
prob = P.laplace_problem(grid)
rng = np.random.default_rng(1); xs = rng.uniform(-1, 1, prob["n"])
b = P.full_from_lower(prob["n"], prob["ptr"], prob["row"], prob["val"]) @ xs
ref = Solver(); ref.analyse(prob["n"], prob["ptr"], prob["row"], prob["order"]); ref.factor(prob["val"]); fr, xr = ref.solve(b)
print("single berr", P.backward_error(prob, xr, b))
er = EmulatedRanks(nr); er.analyse(prob["n"], prob["ptr"], prob["row"], prob["order"])
print("boundaries", er.boundaries)
er.factor(prob["val"])
nt = ref.numeric_tree()
# per-front comparison: which rank holds each node
bad = 0
for i in range(nt["nnodes"]):
    g = ref.node(i)
    holders = []
    for r, s in enumerate(er.s):
        q = s.node(i)
        if q is not None and q["m"] > 0 and q["nelim"] > 0 and q["L"].shape == g["L"].shape:
            d = np.abs(q["L"] - g["L"]).max(); pd = int((q["perm"] != g["perm"]).sum())
            holders.append((r, d, pd))
    if not holders or min(h[1] for h in holders) > 1e-9:
        bad += 1
        if bad < 6: print("node", i, "m,n,ne", g["m"], g["n"], g["nelim"], "holders", holders)
print("bad fronts", bad, "of", nt["nnodes"])
x = er.solve(b)
print("emulated", nr, "ranks: neg", er.num_neg, "moved MB", er.moved / 1e6, "berr", P.backward_error(prob, x, b), "max diff vs single", np.abs(x - xr).max())
