"""Run one BASELINE config end to end on the GPU and print inertia / backward error / timings."""
import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from galahad_b200 import problems as P
from galahad_b200._lib import Solver
cfg = sys.argv[1]
t0 = time.time()
if cfg == "c3":
    prob = P.kkt_grid_problem(110)
elif cfg == "c3s":
    prob = P.kkt_grid_problem(64)
elif cfg == "c4":
    prob = P.qpband_problem(1000000)
elif cfg == "c5":
    prob = P.laplace_problem(int(sys.argv[2]) if len(sys.argv) > 2 else 96, stencil=27)
elif cfg == "c2":
    prob = P.laplace_problem(64)
tg = time.time() - t0
s = Solver()
t0 = time.time(); f = s.analyse(prob["n"], prob["ptr"], prob["row"], prob["order"]); ta = time.time() - t0
inf = s.inform
print(f"{prob['name']}: n={prob['n']} nz={prob['ptr'][-1]} gen {tg:.1f}s analyse {ta:.1f}s flag {f} nodes {inf.num_sup} "
      f"nfact {inf.num_factor:.3e} flops {inf.num_flops:.3e} depth {inf.maxdepth}", flush=True)
nt = s.numeric_tree(); print("  fused nodes", nt["nnodes"], "max front rows", int(np.diff(nt["rptr"]).max()), flush=True)
rng = np.random.default_rng(1)
xs = rng.uniform(-1, 1, prob["n"])
b = P.full_from_lower(prob["n"], prob["ptr"], prob["row"], prob["val"]) @ xs
for rep in range(2):
    t0 = time.time(); f = s.factor(prob["val"]); tw = time.time() - t0
    i = s.inform
    print(f"  factor flag {f} neg {i.num_neg} (expect {prob['inertia'][1]}) two {i.num_two} delay {i.num_delay} nfp {i.not_first_pass} "
          f"nsp {i.not_second_pass} cuda {i.cuda_error} t {i.time_factor*1e3:.1f} ms wall {tw*1e3:.0f} ms  {inf.num_flops/max(i.time_factor,1e-9)/1e12:.2f} TF/s "
          f"launches {i.reserved[0]} factor_bytes {i.factor_bytes/1e9:.2f} GB", flush=True)
    if f < 0: sys.exit(1)
fs, x = s.solve(b)
print(f"  solve flag {fs} t {s.inform.time_solve*1e3:.2f} ms berr {P.backward_error(prob, x, b):.2e} err {np.abs(x-xs).max():.2e}", flush=True)
