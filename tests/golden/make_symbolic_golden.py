"""Writes tests/golden/symbolic_digests.json: per case of tests/test_symbolic_oracle.py the node count, num_factor,
num_flops and a SHA-256 over the integer outputs of the symbolic ORACLE (oracle/symbolic_oracle.py, the restatement of
src/spral/core_analyse.f90 + src/ssids/anal.f90).  python tests/golden/make_symbolic_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
import test_symbolic_oracle as T   # noqa: E402

out = {}
for name, prob, nemin in T.cases():
    o = T.oracle_arrays(prob, nemin)
    out[name] = dict(nemin=nemin, n=int(prob["n"]), nnodes=int(o["nnodes"]), num_factor=int(o["num_factor"]),
                     num_flops=int(o["num_flops"]), sha256=T.digest(o))
json.dump(out, open(os.path.join(HERE, "symbolic_digests.json"), "w"), indent=1, sort_keys=True)
print(len(out), "cases")
