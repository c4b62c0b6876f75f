"""k_front (small fronts in one launch) against the general path: same matrix factorised with B200LDLT_FRONT=0 and =1,
factors compared node by node (bitwise where the algorithms coincide: fronts of at most 32 candidate columns), inertia,
backward error.  GPU only.  python tools/front_check.py [quick]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
from galahad_b200 import problems as P
from galahad_b200._lib import Solver


def run(prob, front, posdef=False):
    os.environ["B200LDLT_FRONT"] = "1" if front else "0"
    s = Solver()
    assert s.analyse(prob["n"], prob["ptr"], prob["row"], prob["order"]) >= 0
    flag = s.factor(prob["val"], posdef)
    nn = s.numeric_tree()["nnodes"]
    nodes = [s.node(i) for i in range(nn)] if flag >= 0 else []
    inf = s.inform
    stats = dict(flag=flag, neg=inf.num_neg, two=inf.num_two, delay=inf.num_delay, nfp=inf.not_first_pass, nsp=inf.not_second_pass,
                 ms=inf.time_factor * 1e3, launches=inf.reserved[0])
    x = berr = None
    if flag >= 0:
        rng = np.random.default_rng(3)
        xs = rng.uniform(-1, 1, prob["n"])
        b = P.full_from_lower(prob["n"], prob["ptr"], prob["row"], prob["val"]) @ xs
        f, x = s.solve(b)
        berr = P.backward_error(prob, x, b)
        # second factorisation: timing with warm buffers
        s.factor(prob["val"], posdef)
        stats["ms2"] = s.inform.time_factor * 1e3
    s.free()
    return stats, nodes, x, berr


def compare(name, prob, posdef=False):
    s0, n0, x0, b0 = run(prob, False, posdef)
    s1, n1, x1, b1 = run(prob, True, posdef)
    ok = s0["flag"] == s1["flag"] and s0["neg"] == s1["neg"] and s0["delay"] == s1["delay"]
    nbit = ndiff = 0
    worst = 0.0
    if s0["flag"] >= 0:
        for a, b in zip(n0, n1):
            if a is None or b is None:
                continue
            same_shape = a["m"] == b["m"] and a["n"] == b["n"] and a["nelim"] == b["nelim"]
            ok = ok and same_shape
            if not same_shape:
                continue
            if np.array_equal(a["L"], b["L"]) and np.array_equal(a["d"][:2 * a["nelim"]], b["d"][:2 * a["nelim"]]) and np.array_equal(a["perm"], b["perm"]):
                nbit += 1
            else:
                ndiff += 1
                if np.array_equal(a["perm"], b["perm"]):
                    worst = max(worst, float(np.abs(a["L"] - b["L"]).max()))
                else:
                    worst = max(worst, float("inf"))
        ok = ok and b1 <= max(1e-14, 10 * b0)
    print(f"{'OK  ' if ok else 'FAIL'} {name}: flag {s0['flag']}/{s1['flag']} neg {s0['neg']}/{s1['neg']} two {s0['two']}/{s1['two']} "
          f"delay {s0['delay']}/{s1['delay']} nfp {s0['nfp']}/{s1['nfp']} nodes bitwise {nbit} differing {ndiff} (max |dL| {worst:.2e}) "
          f"berr {b0}/{b1} ms {s0.get('ms2', 0):.2f}/{s1.get('ms2', 0):.2f} launches {s0['launches']}/{s1['launches']}", flush=True)
    return ok


if __name__ == "__main__":
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    from util import golden_cases
    allok = True
    for name, prob, posdef in golden_cases():
        allok &= compare(name, prob, posdef)
    allok &= compare("lap7_24", P.laplace_problem(24))
    allok &= compare("kkt7_16", P.kkt_grid_problem(16))
    allok &= compare("kkt7_16_hard", P.kkt_grid_problem(16, dlo=1e-10))
    allok &= compare("qpband_50000", P.qpband_problem(50000))
    allok &= compare("qpband_50000_it5", P.qpband_problem(50000, it=5))
    for k in ("indef", "kkt"):
        allok &= compare(f"rand_{k}", P.random_sym_problem(2000, 0.004, k, seed=5))
    allok &= compare("dense_indef_60", P.dense_front_problem(60, "indef")[1])
    allok &= compare("dense_zdiag_40", P.dense_front_problem(40, "zdiag")[1])
    allok &= compare("dense_kkt_64", P.dense_front_problem(64, "kkt")[1])
    allok &= compare("dense_small_30", P.dense_front_problem(30, "small")[1])
    if not quick:
        allok &= compare("lap7_48", P.laplace_problem(48))
        allok &= compare("lap7_64", P.laplace_problem(64))
        allok &= compare("qpband_1000000", P.qpband_problem(1000000))
    print("ALL OK" if allok else "SOME FAILED")
    sys.exit(0 if allok else 1)
