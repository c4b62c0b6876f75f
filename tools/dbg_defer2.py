import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
from galahad_b200 import problems as P
from galahad_b200._lib import Solver
from test_gpu_parity import _arms_problem
A, prob = _arms_problem((300,), 800, "indef")
os.environ["B200LDLT_DEFER_MINM"] = "0"
os.environ["B200LDLT_DEFER"] = sys.argv[1]
os.environ["B200LDLT_DEBUG_UPD"] = "1"
s = Solver()
s.analyse(prob["n"], prob["ptr"], prob["row"], prob["order"])
f = s.factor(prob["val"])
i = s.inform
print("flag", f, "neg", i.num_neg, "expect", prob["inertia"][1], "nfp", i.not_first_pass, flush=True)
