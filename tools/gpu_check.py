"""Development driver: run the CUDA engine against the CPU oracle on a list of problems and print
per-problem diagnostics (inertia, delays, backward error, first differing front).  Used through
gpurun while bringing kernels up; the pytest -m gpu suite is the real gate."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from galahad_b200 import problems as P  # noqa: E402
from galahad_b200._lib import Solver  # noqa: E402
from oracle.refssids import Oracle, RefSSIDS, have_ref, permuted_solve  # noqa: E402


def check(prob, posdef=False, detail=True, use_ref=False, **opts):
    s = Solver(**opts)
    flag = s.analyse(prob["n"], prob["ptr"], prob["row"], prob["order"])
    sym = s.numeric_tree()   # the engine's own (chain-fused) tree: a valid assembly tree for the oracles too
    inf = s.inform
    n = prob["n"]
    rng = np.random.default_rng(1)
    xs = rng.uniform(-1, 1, n)
    A = P.full_from_lower(n, prob["ptr"], prob["row"], prob["val"])
    b = A @ xs
    t = time.time()
    ff = s.factor(prob["val"], posdef)
    tf = time.time() - t
    fi = s.inform.asdict()
    line = (f"{prob['name']:>22s} n={n} nodes={sym['nnodes']} flops={inf.num_flops:.3e} | factor flag={ff} "
            f"neg={fi['num_neg']} two={fi['num_two']} delay={fi['num_delay']} nfp={fi['not_first_pass']} "
            f"nsp={fi['not_second_pass']} cuda={fi['cuda_error']} t={fi['time_factor'] * 1e3:.2f}ms "
            f"(dev {fi['time_factor_device'] * 1e3:.2f}ms, wall {tf * 1e3:.1f}ms) launches={s.inform.reserved[0]}")
    print(line, flush=True)
    ok = ff >= 0
    if ff >= 0:
        fs, x = s.solve(b)
        be = P.backward_error(prob, x, b)
        print(f"{'':>22s} solve flag={fs} berr={be:.2e} err={np.abs(x - xs).max():.2e} "
              f"t={s.inform.time_solve * 1e3:.2f}ms", flush=True)
        ok = ok and fs >= 0 and be < 1e-13
    eng = None
    if use_ref and have_ref():
        eng = RefSSIDS(sym, n, u=s.opt.u, small=s.opt.small_)
        name = "ref"
    elif detail:
        eng = Oracle(sym, n, u=s.opt.u, small=s.opt.small_)
        name = "oracle"
    if eng is not None:
        so = eng.factor(prob["val"], posdef)
        xo = permuted_solve(eng, s.order, b)
        print(f"{'':>22s} {name}: neg={so['num_neg']} two={so['num_two']} delay={so['num_delay']} "
              f"nfp={so['not_first_pass']} nsp={so['not_second_pass']} berr={P.backward_error(prob, xo, b):.2e}", flush=True)
        if ff >= 0 and so["num_neg"] != fi["num_neg"]:
            ok = False
            print(f"{'':>22s} INERTIA MISMATCH", flush=True)
        if detail and name == "oracle" and ff >= 0:
            bad = 0
            for i in range(sym["nnodes"]):
                g, o = s.node(i), eng.node(i)
                if g is None:
                    break
                same = (g["m"], g["n"], g["nelim"]) == (o["m"], o["n"], o["nelim"])
                dl = np.abs(g["L"] - o["L"]).max() if same and g["nelim"] > 0 else 0.0
                if not same or dl > 1e-8 * max(1.0, np.abs(o["L"]).max() if o["L"].size else 1.0):
                    bad += 1
                    if bad <= 3:
                        print(f"{'':>22s} node {i}: gpu (m,n,ne)=({g['m']},{g['n']},{g['nelim']}) "
                              f"oracle=({o['m']},{o['n']},{o['nelim']}) max|dL|={dl:.2e} "
                              f"permdiff={int((g['perm'] != o['perm']).sum()) if same else -1}", flush=True)
            print(f"{'':>22s} fronts differing from oracle: {bad}/{sym['nnodes']}", flush=True)
    s.free()
    return ok


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    allok = True
    if which == "small":
        allok &= check(P.laplace_problem(4))
        allok &= check(P.laplace_problem(8))
        allok &= check(P.laplace_problem(8), posdef=True)
        allok &= check(P.laplace_problem(16))
        allok &= check(P.kkt_grid_problem(8))
        allok &= check(P.kkt_grid_problem(12, dlo=1e-10))
        allok &= check(P.qpband_problem(2000))
        for k in ("spd", "indef", "kkt"):
            allok &= check(P.random_sym_problem(300, 0.03, k))
        for sd in range(3):
            allok &= check(P.random_sym_problem(500, 0.01, "indef", seed=100 + sd))
            allok &= check(P.random_sym_problem(400, 0.02, "kkt", seed=200 + sd))
    elif which == "medium":
        allok &= check(P.laplace_problem(32), detail=False, use_ref=True)
        allok &= check(P.laplace_problem(32), posdef=True, detail=False)
        allok &= check(P.kkt_grid_problem(24), detail=False, use_ref=True)
        allok &= check(P.qpband_problem(100000), detail=False, use_ref=True)
        allok &= check(P.laplace_problem(24, stencil=27), detail=False, use_ref=True)
    elif which == "lap32":
        allok &= check(P.laplace_problem(32))
    elif which == "c2":
        allok &= check(P.laplace_problem(64), detail=False)
        allok &= check(P.laplace_problem(64), posdef=True, detail=False)
    print("ALL OK" if allok else "SOME FAILED", flush=True)
    sys.exit(0 if allok else 1)
