import json
import os

import numpy as np

from galahad_b200 import problems as P

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return json.load(open(os.path.join(GOLD, name)))


def kat_problem(which, order="identity"):
    """The reference's 5x5 systems (src/sls/slst.f90:29-51) as a lower-CSC problem dict."""
    import scipy.sparse as sp
    k = golden("kat_sls.json")
    d = k[which]
    n = k["n"]
    r = np.array(d["row"]) - 1
    c = np.array(d["col"]) - 1
    lo = np.maximum(r, c), np.minimum(r, c)
    A = sp.csc_matrix((np.array(d["val"], float), lo), shape=(n, n))
    A = A + sp.tril(A, -1).T
    N, ptr, row, val = P.lower_csc(A)
    o = np.arange(n, dtype=np.int32) if order == "identity" else (np.array(k["order_reverse"], np.int32) - 1)
    return dict(n=N, ptr=ptr, row=row, val=val, order=o, name=f"kat_{which}_{order}", inertia=None), \
        np.array(d["rhs"], float), np.array(d["sol"], float)


def bcsstk01_problem(order="identity"):
    d = golden("bcsstk01.json")
    n = d["n"]
    o = np.arange(n, dtype=np.int32) if order == "identity" else np.arange(n - 1, -1, -1).astype(np.int32)
    return dict(n=n, ptr=np.array(d["ptr"], np.int64), row=np.array(d["row"], np.int32), val=np.array(d["val"], float),
                order=o, name=f"bcsstk01_{order}", inertia=(n, 0, 0))


def golden_cases():
    """(name, problem, posdef) exactly as tests/golden/make_golden.py generated them."""
    cases = [("lap7_8", P.laplace_problem(8), False), ("lap7_8_pd", P.laplace_problem(8), True),
             ("lap7_12", P.laplace_problem(12), False), ("lap27_10", P.laplace_problem(10, stencil=27), False),
             ("kkt7_8", P.kkt_grid_problem(8), False), ("kkt7_12_hard", P.kkt_grid_problem(12, dlo=1e-10), False),
             ("qpband_2000", P.qpband_problem(2000), False), ("qpband_2000_it3", P.qpband_problem(2000, it=3), False)]
    for k in ("spd", "indef", "kkt"):
        for sd in range(3):
            cases.append((f"rand_{k}_{sd}", P.random_sym_problem(300 + 50 * sd, 0.02, k, seed=1000 + sd), False))
    return cases


def rhs_for(prob, seed=7):
    rng = np.random.default_rng(seed)
    xs = rng.uniform(-1, 1, prob["n"])
    A = P.full_from_lower(prob["n"], prob["ptr"], prob["row"], prob["val"])
    return xs, A @ xs
