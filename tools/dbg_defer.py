import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
from galahad_b200 import problems as P
from galahad_b200._lib import Solver
from test_gpu_parity import _arms_problem
cases = [((300,), 800, "indef"), ((520, 300), 1000, "indef"), ((640,), 1100, "indef")]
for arms, ns, kind in cases:
    A, prob = _arms_problem(arms, ns, kind)
    for env in ({"B200LDLT_DEFER": "1"}, {"B200LDLT_DEFER": "2"}, {"B200LDLT_DEFER": "2", "B200LDLT_LOOKAHEAD": "0"}, {"B200LDLT_DEFER": "4"}, {"B200LDLT_UPD3": "0"}):
        for k in ("B200LDLT_DEFER", "B200LDLT_LOOKAHEAD", "B200LDLT_UPD3"): os.environ.pop(k, None)
        os.environ["B200LDLT_DEFER_MINM"] = "0"
        os.environ.update(env)
        s = Solver()
        s.analyse(prob["n"], prob["ptr"], prob["row"], prob["order"])
        f = s.factor(prob["val"])
        i = s.inform
        b = A @ np.ones(prob["n"])
        fs, x = s.solve(b)
        be = np.abs(A @ x - b).max() / (np.abs(A).sum(1).max() * np.abs(x).max() + np.abs(b).max())
        print(arms, kind, env, "flag", f, "neg", i.num_neg, "expect", prob["inertia"][1], "two", i.num_two, "delay", i.num_delay, "nfp", i.not_first_pass, "nsp", i.not_second_pass, "berr %.2e" % be, flush=True)
