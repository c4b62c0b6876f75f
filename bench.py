#!/usr/bin/env python
"""bench.py -- SLS factorize (FP64 TFLOP/s) + solve (ms) on BASELINE.json config C2
(3-D 7-point Laplacian 64^3, 262,144 unknowns, geometric nested-dissection order, SLS default
pivot_control=1 i.e. the pivoted LDL^T path) through the b200ldlt C ABI.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--grid G]

A step = one numeric factorization + one solve (1 RHS) of the same matrix (analyse once, as the
interior-point callers do).  `value` = num_flops / mean factorize time with A's values resident in
HBM (num_flops = the reference's own count, core_analyse.f90:880-892); `solve_ms` rides along.
`e2e` = the same through the host-buffer C-ABI calls (H2D of val and b, D2H of x inside the timed
region).  N > 1: ONE factorisation of one matrix, tree-parallel over the ranks (subtrees per GPU,
contribution blocks shipped to the owner of the parent front with NCCL).  The
default is WEAK scaling: the grid grows with N so that the factorisation's flops per GPU stay those of
C2 (64^3 at N = 1; 72^3, 80^3, 90^3 at N = 2, 4, 8: 2.02, 3.83, 7.86 times the flops) -- a 35 ms
problem cannot occupy eight GPUs, the large problems the partition exists for can.  `--scaling strong`
keeps the 64^3 matrix at every N.
`--impl reference` times the reference's own C++ SSIDS CPU engine (oracle/_ref) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum of one k_upd3 launch from an ncu --set full capture (profiles/); filled in
# from the capture of the current build, None until then
UPD3_TRAFFIC = {"bytes": 2453002600,
                "note": "dram__bytes_read.sum + dram__bytes_write.sum of the longest k_upd3 launch of C2 (1.71 ms, rank-512 deferred update of a "
                        "level's fronts, ncu --set full, profiles/r2_ncu_k_upd3.txt): 2.09 GB read + 0.36 GB written for about 0.8 GB of "
                        "algorithmic bytes -- the operands of all fronts of the launch (about 200 MB) do not stay in L2 beside the C stream; the "
                        "launch is compute-bound all the same (DMMA pipe 88 % active, 1.4 TB/s of DRAM traffic); one launch, not an average"}

METRIC = "sls_factorize_fp64_tflops"
UNIT = "TFLOP/s"


# weak scaling: grid with N times the factorisation flops of 64^3 (flops ~ g^6; engine-tree counts 8.875e11, 1.681e12,
# 3.452e12 against 4.391e11)
WEAK_GRID = {1: 64, 2: 72, 4: 80, 8: 90}


def grid_for(args):
    if args.grid:
        return args.grid
    if args.scaling == "strong" or args.gpus <= 1:
        return 64
    return WEAK_GRID.get(args.gpus, int(round(64 * args.gpus ** (1.0 / 6.0))))


def make_problem(grid):
    from galahad_b200 import problems as P
    return P.laplace_problem(grid)


def config_for(grid, flops, nfact):
    """The `config` object, identical in both arms (the driver compares them)."""
    return {"workload": workload_name(grid), "num_flops": flops, "num_factor": nfact}


def workload_name(grid):
    tag = "C2" if grid == 64 else "C2 scaled to N GPUs (flops per GPU as 64^3)"
    return f"{tag}: 3-D 7-point Laplacian {grid}^3 (n={grid ** 3}), geometric ND order, SLS pivot_control=1 (LDL^T, u=0.01), factor + 1-RHS solve"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def mark(self):
        """Start of the timed region: earlier samples (warm-up) are only used if nvidia-smi delivers
        none inside the region (it samples every 100 ms and needs about a second to start)."""
        self.t_mark = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t_mark = getattr(self, "t_mark", 0.0)
        inside = [ln for (t, ln) in self.lines if t >= t_mark]
        window = "timed region"
        if len(inside) < 2:
            inside = [ln for (t, ln) in self.lines]
            window = "warm-up + timed region (GPU under the same load)"
        for ln in inside:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        hi = [s for s in sm if s > 0.5 * max(sm)]
        return {"sm_mhz": float(np.median(hi if hi else sm)), "sm_max_mhz": float(max(mx)),
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def solve_roofline(gbs):
    """The solve against the HBM copy rate: MEASURED_PEAKS.json (driver-written) if present, else the profiling recipe's fallback."""
    peak, src = 6650.0, "fallback of B200_PROFILING.md (no MEASURED_PEAKS.json)"
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")) as fh:
            peak, src = float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    return {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "peak_source": src,
            "bytes": "2 sweeps x 8 B x (entries of L + n): every entry of the factors read once per sweep"}


def measure_fp64_peak(torch):
    """cuBLAS Dgemm through torch (float64 4096^3, best of 5): the FP64 denominator.  The driver's
    MEASURED_PEAKS.json has no FP64 entry, so it is measured here, same method as its bf16 entry."""
    n = 4096
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b, c
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import refssids
    if not refssids.have_ref():
        try:
            refssids.build_ref()
        except Exception as e:  # the reference cannot be built on this box
            print(json.dumps({"impl": "reference", "unavailable": f"oracle/_ref missing and not buildable: {e}"}))
            return 0
    from galahad_b200._lib import Solver
    cores = os.cpu_count() or 1
    # bounded sample: shrink the grid until (K+W) steps fit in ~3 minutes (flops ~ g^6, measured
    # 0.035 TFLOP/s on 8 cores for g=64)
    total_steps = args.steps + args.warmup
    grid = args.grid
    est_rate = 0.0045e12 * min(cores, 32)
    for g in (args.grid, 80, 72, 64, 56, 48, 40, 32, 24):
        if g > args.grid:
            continue
        grid = g
        est = 4.34e11 * (g / 64.0) ** 6 / est_rate
        if est * total_steps <= 180.0:
            break
    prob = make_problem(grid)
    s = Solver()
    s.analyse(prob["n"], prob["ptr"], prob["row"], prob["order"])
    flops = float(s.inform.num_flops)
    cfg_flops, cfg_nfact = flops, int(s.inform.num_factor)
    if grid != args.grid:
        # the config names the full workload; the timed sample is the smaller grid
        s2 = Solver()
        p2 = make_problem(args.grid)
        s2.analyse(p2["n"], p2["ptr"], p2["row"], p2["order"])
        cfg_flops, cfg_nfact = float(s2.inform.num_flops), int(s2.inform.num_factor)
        s2.free()
        del p2
    sym = s.symbolic()
    ref = refssids.RefSSIDS(sym, prob["n"], u=0.01, small=1e-20)
    rng = np.random.default_rng(20261018)
    b = rng.uniform(-1, 1, prob["n"])
    tf, ts = [], []
    for it in range(total_steps):
        st = ref.factor(prob["val"], False)
        t0 = time.perf_counter()
        refssids.permuted_solve(ref, s.order, b)
        t1 = time.perf_counter() - t0
        if it >= args.warmup:
            tf.append(st["seconds"]); ts.append(t1)
    mf, ms = float(np.mean(tf)), float(np.mean(ts))
    val = flops / mf / 1e12
    sample = f"full factor+solve of the {grid}^3 grid ({flops:.3e} flops) per step, {len(tf)} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": (mf + ms) * 1e3, "higher_is_better": True, "scaling": args.scaling_kind,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_for(args.grid, cfg_flops, cfg_nfact),
        "sample_grid": grid, "threads": ref.nthreads(),
        "factor_ms": mf * 1e3, "solve_ms": ms * 1e3,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": ref.nthreads(), "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from galahad_b200._lib import Solver
    prob = make_problem(args.grid)
    n = prob["n"]
    if world > 1:
        return run_tree_parallel(args, prob, torch, dist, rank, world, local)
    s = Solver(device=local)
    s.analyse(n, prob["ptr"], prob["row"], prob["order"])
    flops = float(s.inform.num_flops)
    nfact = int(s.inform.num_factor)
    rng = np.random.default_rng(20261018)
    b = rng.uniform(-1, 1, n)
    d_val = torch.from_numpy(prob["val"]).cuda()
    d_b = torch.from_numpy(b).cuda()
    d_x = torch.empty_like(d_b)
    h_val = torch.from_numpy(prob["val"]).pin_memory()
    h_b = torch.from_numpy(b.copy()).pin_memory()
    h_x = torch.empty(n, dtype=torch.float64).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        flush.zero_()                      # > L2 (126 MB) written between iterations
        d_x.copy_(d_b)
        torch.cuda.synchronize()
        f = s.factor_device(d_val.data_ptr(), False)
        assert f >= 0, f"factor flag {f}"
        tf = s.inform.time_factor_device
        nl = s.inform.reserved[0]
        uf, us, ul = s.inform.upd_flops, s.inform.upd_seconds, s.inform.reserved[2]
        f = s.solve_device(d_x.data_ptr(), 1, n)
        assert f >= 0, f"solve flag {f}"
        return tf, s.inform.time_solve, nl + s.inform.reserved[1], uf, us, ul

    def step_host():
        import ctypes as C
        flush.zero_()
        h_x.copy_(h_b)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        f = s.L.b200ldlt_factor(s.akeep, C.byref(s.fkeep), 0, C.cast(h_val.data_ptr(), C.POINTER(C.c_double)),
                                C.byref(s.opt), C.byref(s.inform))
        t1 = time.perf_counter()
        assert f >= 0, f
        f = s.L.b200ldlt_solve(s.akeep, s.fkeep, 0, 1, C.cast(h_x.data_ptr(), C.POINTER(C.c_double)), n,
                               C.byref(s.opt), C.byref(s.inform))
        t2 = time.perf_counter()
        assert f >= 0, f
        return t1 - t0, t2 - t1

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = [step_device() for _ in range(args.steps)]
    e1.record()
    barrier()
    total_ms = e0.elapsed_time(e1)
    tf = float(np.mean([r[0] for r in res]))
    ts = float(np.mean([r[1] for r in res]))
    launches = int(np.sum([r[2] for r in res]))
    # roofline pass: the same K steps with CUDA events around every launch of the dominant kernel
    # (profile mode keeps the factorisation on one stream so that a launch is timed alone)
    s.opt.reserved[0] = 2
    step_device()
    barrier()
    resp = [step_device() for _ in range(args.steps)]
    barrier()
    s.opt.reserved[0] = 0
    clocks = sampler.stop() if rank == 0 else None
    upd_flops = float(np.sum([r[3] for r in resp]))
    upd_secs = float(np.sum([r[4] for r in resp]))
    tf_prof = float(np.mean([r[0] for r in resp]))
    # check the answer of the last step (not timed)
    x = d_x.cpu().numpy()
    from galahad_b200 import problems as P
    berr = P.backward_error(prob, x, b)
    # eight right-hand sides, resident (SURVEY section 8d ii): columns of the same b, after the last factorisation
    d_x8 = d_b.repeat(8).contiguous()
    ts8 = []
    for _ in range(3):
        d_x8.copy_(d_b.repeat(8))
        torch.cuda.synchronize()
        f = s.solve_device(d_x8.data_ptr(), 8, n)
        assert f >= 0, f"solve flag {f}"
        ts8.append(s.inform.time_solve)
    ts8 = float(np.mean(ts8[1:]))
    assert torch.equal(d_x8[:n], d_x8[7 * n:]) and torch.equal(d_x8[:n].cpu(), torch.from_numpy(x)), "columns of a multi-column solve differ"
    del d_x8
    # end to end through the host-buffer C ABI
    for _ in range(2):
        step_host()
    barrier()
    eh = [step_host() for _ in range(args.steps)]
    barrier()
    ef = float(np.mean([r[0] for r in eh]))
    es = float(np.mean([r[1] for r in eh]))
    # max over ranks
    t = torch.tensor([tf, ts, total_ms, ef, es], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tf, ts, total_ms, ef, es = [float(v) for v in t.cpu()]
    if rank == 0:
        peak = measure_fp64_peak(torch)
        value = world * flops / tf / 1e12
        achieved = upd_flops / upd_secs / 1e12 if upd_secs > 0 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling_kind, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_for(args.grid, flops, nfact),
            "notes": {"parallelism": "1 GPU",
                      "l2": "256 MB buffer written between iterations; factor working set >> L2",
                      "timing": "value = flops / mean CUDA-event time of the numeric factorization (engine stream, val resident); ms_per_step = CUDA events around the K steps / K"},
            "factor_ms": tf * 1e3, "solve_ms": ts * 1e3,
            "solve_gbs": (2 * 8 * nfact + 2 * 8 * n) / ts / 1e9,
            "solve_8rhs_ms": ts8 * 1e3,
            "solve_roofline": solve_roofline((2 * 8 * nfact + 2 * 8 * n) / ts / 1e9),
            "backward_error": berr,
            "gpu_launches": launches,
            "e2e": {"value": world * flops / ef / 1e12, "unit": UNIT, "h2d_bytes_per_step": int(8 * prob["val"].size + 8 * n),
                    "d2h_bytes_per_step": int(8 * n), "factor_ms": ef * 1e3, "solve_ms": es * 1e3},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": (achieved / peak) if achieved else None,
                         "traffic": UPD3_TRAFFIC.get("bytes") if args.grid == 64 else None,
                         "traffic_note": UPD3_TRAFFIC.get("note"),
                         "kernel": "k_upd3 (trailing Schur update of the large fronts, FP64 DMMA)",
                         "peak_source": "measured here: cuBLAS Dgemm 4096^3 via torch.matmul(float64), best of 5 (MEASURED_PEAKS.json has no FP64 entry)",
                         "share_of_factor_time": upd_secs / (tf_prof * args.steps) if tf_prof > 0 else None,
                         "launches_per_step": int(resp[-1][5]),
                         "measured": "second pass of the same K steps with CUDA events around every k_upd3 launch on the engine stream (no look-ahead: each launch timed alone, factor_ms %.2f)" % (tf_prof * 1e3)},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(prob, s, flops)
        if world == 1 and not args.no_kkt:
            s.free()
            del d_val, flush
            torch.cuda.empty_cache()
            line["kkt_c3"] = kkt_c3(torch, peak)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_tree_parallel(args, prob, torch, dist, rank, world, local):
    """N > 1: one matrix, subtree-partitioned across the ranks (galahad_b200/dist.py)."""
    import numpy as np
    from galahad_b200 import problems as P
    from galahad_b200.dist import DistSolver
    n = prob["n"]
    ds = DistSolver(device=local)
    ds.analyse(n, prob["ptr"], prob["row"], prob["order"])
    flops = float(ds.s.inform.num_flops)
    nfact = int(ds.s.inform.num_factor)
    rng = np.random.default_rng(20261018)
    b = rng.uniform(-1, 1, n)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def step():
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        f = ds.factor(prob["val"])
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        assert f >= 0, f
        x = ds.solve(b)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1, x

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = [step() for _ in range(args.steps)]
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    tf = float(np.mean([r[0] for r in res]))
    ts = float(np.mean([r[1] for r in res]))
    t = torch.tensor([tf, ts, e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tf, ts, total_ms = [float(v) for v in t.cpu()]
    if rank == 0:
        x = res[-1][2]
        line = {
            "metric": METRIC, "value": flops / tf / 1e12, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling_kind, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_for(args.grid, flops, nfact),
            "notes": {"parallelism": f"tree-parallel over {world} GPUs: proportional mapping, {ds.nstages} stages, "
                                     f"{sum(1 for st in ds.boundaries for (_, a, c) in st if a != c)} boundary fronts moved between ranks "
                                     f"({ds.bytes_moved / 1e6:.1f} MB through rank 0 per factorisation)",
                      "l2": "256 MB buffer written between iterations; factor working set >> L2",
                      "timing": "max over ranks of the wall time of factor (H2D of val, the stages and the exchanges between them) between device synchronisations"},
            "factor_ms": tf * 1e3, "solve_ms": ts * 1e3, "stage_ms_rank0": [round(v * 1e3, 3) for v in ds.t_stage],
            "backward_error": P.backward_error(prob, x, b), "num_neg": ds.num_neg,
            "gpu_launches": int(ds.launches) * args.steps,   # rank 0, all stages
            "e2e": {"value": flops / tf / 1e12, "unit": UNIT, "h2d_bytes_per_step": int(world * 8 * prob["val"].size + world * 8 * n),
                    "d2h_bytes_per_step": int(8 * n), "factor_ms": tf * 1e3, "solve_ms": ts * 1e3},
            "roofline": {"bound": "tensor", "achieved": None, "peak": None, "unit": "TFLOP/s", "frac": None, "traffic": None,
                         "kernel": "k_update (see the N=1 line; per-launch profiling is off in the multi-rank run)"},
            "clocks": clocks,
        }
        print(json.dumps(line))
    dist.destroy_process_group()
    return 0


def kkt_c3(torch, peak):
    """BASELINE config C3 beside the headline: the indefinite KKT matrix [H A^T; A -D] of a 110^3 grid QP (n = 1,996,500,
    1.04e14 flops), where the trailing updates, not the panel chain, bound the factorisation.  One warm-up and one timed
    factorisation + solve (CUDA events on the engine stream, values resident); inertia checked against (n, m, 0)."""
    try:
        from galahad_b200 import problems as P
        from galahad_b200._lib import Solver
        prob = P.kkt_grid_problem(110)
        s = Solver()
        s.analyse(prob["n"], prob["ptr"], prob["row"], prob["order"])
        flops = float(s.inform.num_flops)
        d_val = torch.from_numpy(prob["val"]).cuda()
        rng = np.random.default_rng(20261018)
        b = rng.uniform(-1, 1, prob["n"])
        d_x = torch.from_numpy(b).cuda()
        tf = 0.0
        for _ in range(2):
            torch.cuda.synchronize()
            f = s.factor_device(d_val.data_ptr(), False)
            assert f >= 0, f
            tf = s.inform.time_factor_device
        neg = int(s.inform.num_neg)
        f = s.solve_device(d_x.data_ptr(), 1, prob["n"])
        assert f >= 0, f
        ts = s.inform.time_solve
        berr = P.backward_error(prob, d_x.cpu().numpy(), b)
        out = {"workload": "C3: KKT [H A^T; A -D] of a 110^3 grid QP (n=%d), ND order, pivot_control=1" % prob["n"],
               "num_flops": flops, "factor_ms": tf * 1e3, "tflops": flops / tf / 1e12, "frac_of_fp64_peak": flops / tf / 1e12 / peak,
               "solve_ms": ts * 1e3, "solve_gbs": (2 * 8 * int(s.inform.num_factor)) / ts / 1e9,
               "num_neg": neg, "num_neg_expected": int(prob["inertia"][1]), "num_delay": int(s.inform.num_delay),
               "backward_error": berr, "steps": 1, "warmup": 1}
        s.free()
        return out
    except Exception as e:   # the headline line must survive a failure of the secondary measurement
        return {"unavailable": repr(e)}


def cpu_baseline(prob, s, flops):
    """The reference's own CPU engine (oracle/_ref), one full factorization of the same matrix on
    the box's host cores.  Reported context, not the target."""
    try:
        from oracle import refssids
        if not refssids.have_ref():
            refssids.build_ref()
        sym = s.symbolic()
        ref = refssids.RefSSIDS(sym, prob["n"], u=0.01, small=1e-20)
        st = ref.factor(prob["val"], False)
        out = {"value": flops / st["seconds"] / 1e12, "unit": UNIT, "cores": ref.nthreads(), "kind": "reference",
               "sample": f"one full numeric factorization of the same matrix ({flops:.3e} flops, {st['seconds']:.2f} s)",
               "num_neg": st["num_neg"]}
        ref.close()
        return out
    except Exception as e:
        return {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"unavailable: {e}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200ldlt")
    ap.add_argument("--grid", type=int, default=0, help="grid size (default: 64 at N = 1, scaled with N otherwise)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = flops per GPU as C2 (default), strong = the 64^3 matrix at every N")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-kkt", action="store_true", help="skip the secondary C3 (KKT) measurement")
    args = ap.parse_args()
    args.scaling_kind = args.scaling if args.gpus > 1 else "weak"
    args.grid = grid_for(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
