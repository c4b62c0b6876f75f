"""Flops, fronts and shapes per level of the engine tree (CPU only)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from galahad_b200 import problems as P
from galahad_b200._lib import Solver
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 64
stencil = int(sys.argv[2]) if len(sys.argv) > 2 else 7
prob = P.laplace_problem(grid, stencil=stencil) if stencil in (7, 27) else P.kkt_grid_problem(grid)
s = Solver()
s.analyse(prob["n"], prob["ptr"], prob["row"], prob["order"])
nt = s.numeric_tree()
nn = nt["nnodes"]; sptr = nt["sptr"]; rptr = nt["rptr"]; par = nt["sparent"]
ncol = np.diff(sptr); nrow = np.diff(rptr)
# level = height above the deepest leaf as the engine does: depth levels, deepest first
depth = np.zeros(nn, int)
for i in range(nn - 1, -1, -1):
    p = par[i]
    depth[i] = depth[p] + 1 if p < nn else 0
maxd = depth.max()
lvl = maxd - depth
fl = np.zeros(nn)
for i in range(nn):
    m, k = float(nrow[i]), float(ncol[i])
    j = np.arange(m - k + 1, m + 1)
    fl[i] = (j * j).sum()
tot = fl.sum()
print(f"n={prob['n']} nodes={nn} flops={tot:.3e}")
for l in range(maxd + 1):
    idx = np.where(lvl == l)[0]
    print(f"level {l:2d} fronts {len(idx):5d} flops {fl[idx].sum():.3e} ({fl[idx].sum()/tot*100:5.1f}%)  at 30TF {fl[idx].sum()/30e12*1e3:6.3f} ms  "
          f"ncol max {ncol[idx].max():5d} nrow max {nrow[idx].max():5d} mean ncol {ncol[idx].mean():7.1f} nrow {nrow[idx].mean():7.1f}")
