"""Generates tests/golden/*.json.  Run HERE (needs /root/reference and oracle/_ref); the fixtures it
writes are committed so that the test-suite never reads /root/reference.

  kat_sls.json     the reference's own known-answer systems (src/sls/slst.f90:29-51, same data in
                   src/sls/C/slst.c:23-30 and src/ssids/ssidss.data) -- copied values, 5x5.
  bcsstk01.json    examples/bcsstk01.tar.gz (Rutherford-Boeing, n=48, 224 entries) converted to
                   0-based lower CSC; driver semantics x=1, b=Ax (src/sls/runsls_rb.f90:126-137).
  ref_runs.json    outputs of the reference's C++ SSIDS engine (oracle/_ref) on seeded inputs:
                   inertia / delays / solution digests, used to pin oracle/ldlt_oracle.c and the
                   CUDA engine on boxes where the reference cannot be rebuilt.
"""
import io
import json
import os
import sys
import tarfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, ROOT)
REF = "/root/reference"


def fortran_ints(lines, count, width):
    out = []
    for ln in lines:
        for i in range(0, len(ln.rstrip("\n")), width):
            tok = ln[i:i + width].strip()
            if tok:
                out.append(int(tok))
    assert len(out) >= count
    return out[:count]


def read_rb(text):
    lines = text.splitlines()
    l2 = lines[1].split()
    ptrcrd, indcrd, valcrd = int(l2[1]), int(l2[2]), int(l2[3])
    l3 = lines[2].split()
    n, nz = int(l3[1]), int(l3[3])
    body = lines[4:]
    ptr = fortran_ints(body[:ptrcrd], n + 1, 4)
    ind = fortran_ints(body[ptrcrd:ptrcrd + indcrd], nz, 3)
    vals = []
    for ln in body[ptrcrd + indcrd:ptrcrd + indcrd + valcrd]:
        for i in range(0, len(ln), 20):
            tok = ln[i:i + 20].strip()
            if tok:
                vals.append(float(tok.replace("D", "E")))
    return n, np.array(ptr) - 1, np.array(ind) - 1, np.array(vals[:nz])


def main():
    # --- KATs (values as printed in the reference test, 1-based -> kept 1-based here) ---
    kat = {
        "source": "src/sls/slst.f90:29-51",
        "n": 5,
        "indef": {"row": [1, 2, 2, 3, 3, 4, 5], "col": [1, 1, 5, 2, 3, 3, 5], "val": [2, 3, 6, 4, 1, 5, 1],
                  "rhs": [8, 45, 31, 15, 17], "sol": [1, 2, 3, 4, 5]},
        "spd": {"row": [1, 2, 3, 3, 4, 5, 5], "col": [1, 2, 2, 3, 4, 1, 5], "val": [6, 7, 2, 3, 4, 1, 5],
                "rhs": [11, 20, 13, 16, 26], "sol": [1, 2, 3, 4, 5]},
        "order_reverse": [5, 4, 3, 2, 1],
        "ssidss_data": {"source": "src/ssids/ssidss.data", "ptr": [1, 3, 6, 8, 9, 10], "row": [1, 2, 2, 3, 5, 3, 4, 4, 5],
                        "val": [2.0, 3.0, 0.0, 4.0, 6.0, 1.0, 5.0, 0.0, 1.0], "rhs": [8, 45, 31, 15, 17]},
    }
    json.dump(kat, open(os.path.join(HERE, "kat_sls.json"), "w"), indent=1)

    # --- bcsstk01 ---
    with tarfile.open(os.path.join(REF, "examples", "bcsstk01.tar.gz")) as tf:
        text = tf.extractfile("bcsstk01/bcsstk01.rb").read().decode()
    n, ptr, row, val = read_rb(text)
    json.dump({"source": "examples/bcsstk01.tar.gz (bcsstk01.rb)", "n": n, "ptr": ptr.tolist(), "row": row.tolist(),
               "val": val.tolist()}, open(os.path.join(HERE, "bcsstk01.json"), "w"))

    # --- reference-engine runs on seeded inputs ---
    from galahad_b200 import problems as P
    from galahad_b200._lib import Solver
    from oracle.refssids import RefSSIDS, permuted_solve
    runs = []
    cases = [("lap7_8", P.laplace_problem(8), False), ("lap7_8_pd", P.laplace_problem(8), True),
             ("lap7_12", P.laplace_problem(12), False), ("lap27_10", P.laplace_problem(10, stencil=27), False),
             ("kkt7_8", P.kkt_grid_problem(8), False), ("kkt7_12_hard", P.kkt_grid_problem(12, dlo=1e-10), False),
             ("qpband_2000", P.qpband_problem(2000), False), ("qpband_2000_it3", P.qpband_problem(2000, it=3), False)]
    for k in ("spd", "indef", "kkt"):
        for sd in range(3):
            cases.append((f"rand_{k}_{sd}", P.random_sym_problem(300 + 50 * sd, 0.02, k, seed=1000 + sd), False))
    for name, prob, posdef in cases:
        s = Solver()
        s.analyse(prob["n"], prob["ptr"], prob["row"], prob["order"])
        sym = s.symbolic()
        ref = RefSSIDS(sym, prob["n"])
        st = ref.factor(prob["val"], posdef, nthreads=1)
        rng = np.random.default_rng(7)
        xs = rng.uniform(-1, 1, prob["n"])
        A = P.full_from_lower(prob["n"], prob["ptr"], prob["row"], prob["val"])
        b = A @ xs
        x = permuted_solve(ref, s.order, b)
        runs.append({"name": name, "posdef": posdef, "n": prob["n"], "num_sup": int(sym["nnodes"]),
                     "num_factor": int(s.inform.num_factor), "num_flops": int(s.inform.num_flops),
                     "flag": st["flag"], "num_neg": st["num_neg"], "num_zero": st["num_zero"],
                     "num_delay": st["num_delay"], "berr": P.backward_error(prob, x, b),
                     "x_head": x[:8].tolist(), "x_sum": float(x.sum())})
        ref.close(); s.free()
    json.dump({"engine": "reference C++ SSIDS (src/ssids/cpu), pivot_method=app_block, u=0.01, small=1e-20, 1 thread",
               "runs": runs}, open(os.path.join(HERE, "ref_runs.json"), "w"), indent=1)
    print("golden fixtures written:", os.listdir(HERE))


if __name__ == "__main__":
    main()
