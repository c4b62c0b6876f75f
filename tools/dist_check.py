"""Multi-GPU check: torchrun --nproc-per-node N tools/dist_check.py [grid]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, torch.distributed as dist
from galahad_b200 import problems as P
from galahad_b200.dist import DistSolver
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29511")
if "RANK" in os.environ:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank() if dist.is_initialized() else 0
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 32
kind = sys.argv[2] if len(sys.argv) > 2 else "lap"
prob = P.laplace_problem(grid) if kind == "lap" else P.laplace_problem(grid, stencil=27) if kind == "lap27" else P.kkt_grid_problem(grid)
ds = DistSolver()
ds.debug = bool(os.environ.get('B200LDLT_DIST_DEBUG'))
ds.analyse(prob["n"], prob["ptr"], prob["row"], prob["order"])
rng = np.random.default_rng(1); xs = rng.uniform(-1, 1, prob["n"])
b = P.full_from_lower(prob["n"], prob["ptr"], prob["row"], prob["val"]) @ xs
for rep in range(3):
    torch.cuda.synchronize(); 
    if dist.is_initialized(): dist.barrier()
    t0 = time.time(); flag = ds.factor(prob["val"]); torch.cuda.synchronize(); tf = time.time() - t0
    if flag < 0:
        print(f"[rank {rank} of {ds.world}] factor flag {flag} own flag {ds.s.inform.flag} cuda_error {ds.s.inform.cuda_error} stages done {len(ds.t_stage)} "
              f"stage times {[round(v*1e3,1) for v in ds.t_stage]} free GB {torch.cuda.mem_get_info()[0]/1e9:.1f}", flush=True)
        break
    t0 = time.time(); x = ds.solve(b); ts = time.time() - t0
    print(f"[rank {rank} of {ds.world}] rep {rep} {prob['name']} flag {flag} neg {ds.num_neg} (expect {prob['inertia'][1]}) delay {ds.num_delay} "
          f"stages {ds.nstages} moved {ds.bytes_moved/1e6:.1f} MB factor wall {tf*1e3:.1f} ms "
          f"(stages {[round(v*1e3,1) for v in ds.t_stage]}) solve wall {ts*1e3:.1f} ms berr {P.backward_error(prob, x, b):.2e} sum|x| {np.abs(x).sum():.6e}", flush=True)
if dist.is_initialized(): dist.destroy_process_group()
