"""Load balance of the proportional mapping (CPU only): flops per rank and stage.  python tools/map_balance.py grid stencil nranks"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from galahad_b200 import problems as P
from galahad_b200._lib import Solver
grid = int(sys.argv[1]); stencil = int(sys.argv[2]); nr = int(sys.argv[3])
prob = P.laplace_problem(grid, stencil=stencil)
s = Solver(); s.analyse(prob["n"], prob["ptr"], prob["row"], prob["order"])
nt = s.numeric_tree(); nn = nt["nnodes"]
ncol = np.diff(nt["sptr"]).astype(float); nrow = np.diff(nt["rptr"]).astype(float)
m = nrow - ncol
fl = ncol * m * m + ncol * ncol * m + ncol ** 3 / 3
s.L.b200ldlt_partition(s.akeep, nr)
own = np.zeros(nn, int); stg = np.zeros(nn, int)
for i in range(nn):
    a, b = C.c_int(), C.c_int()
    s.L.b200ldlt_map_info(s.akeep, i, C.byref(a), C.byref(b)); own[i] = a.value; stg[i] = b.value
tot = fl.sum()
print(f"{prob['name']} nranks {nr} stages {stg.max()+1} total flops {tot:.3e}")
crit = 0.0
for st in range(stg.max() + 1):
    w = np.array([fl[(own == r) & (stg == st)].sum() for r in range(nr)])
    crit += w.max()
    print(f" stage {st}: per-rank share of total (%):", " ".join(f"{100*v/tot:5.1f}" for v in w), f" max {100*w.max()/tot:.1f}")
print(f" critical path {100*crit/tot:.1f} % of the flops -> speed-up bound {tot/crit:.2f}")
