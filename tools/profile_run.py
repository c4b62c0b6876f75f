"""Short run for ncu: analyse once, factor+solve `reps` times on the C2 grid (or --grid G)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from galahad_b200 import problems as P
from galahad_b200._lib import Solver
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
prob = P.laplace_problem(grid)
s = Solver()
s.analyse(prob["n"], prob["ptr"], prob["row"], prob["order"])
b = np.ones(prob["n"])
for r in range(reps):
    f = s.factor(prob["val"])
    assert f >= 0, f
    print("factor ms", s.inform.time_factor * 1e3, "launches", s.inform.reserved[0], flush=True)
    f, x = s.solve(b)
    print("solve ms", s.inform.time_solve * 1e3, "launches", s.inform.reserved[1], flush=True)
print("berr", P.backward_error(prob, x, b))
