#!/usr/bin/env python
"""bench.py -- headline benchmark of the faer-amg hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--grid 256]

Workload (BASELINE.json north_star / configs[3]): PCG + AMG V(1,1)-cycle solve of the 3-D 7-point
Poisson problem on an n^3 grid (default 256^3, 16.8 M unknowns) to rel. residual 1e-8, b = 1, zero
initial guess; smoothed-aggregation hierarchy over 2x2x2 geometric aggregates built on the GPU
(SpGEMM RAP), L1-Jacobi smoother, exact coarsest solve.  One *step* = one full solve.  The same
problem is solved at every N ("strong" scaling); at N > 1 the fine levels are row-partitioned with
NCCL halo exchange.

metric  = PCG+AMG solve time (ms, lower is better), device-resident vectors          -> "value"
e2e     = the same solve through the host-pointer C ABI call (pinned host b and x,
          H2D + D2H inside the timed region)                                         -> "e2e"
roofline= the dominant kernel (fused diagonal smoother sweep on the fine level; SpMV and residual
          reported next to it): algorithmic bytes / CUDA-event time vs the measured HBM copy peak
cpu_baseline / --impl reference = the CPU restatement of the reference (oracle/, "port"; the Rust
          crate cannot be built here) on the box's host cores.
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

REL_TOL = 1e-8


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""

    def __init__(self, device=0):
        self.rows, self.proc, self.device = [], None, device

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def golden_iters(n):
    try:
        g = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_golden.json")))["amg"]
        return int(g[f"g7_{n}_l1"]["iters"]["1e-08"]["iters"])
    except Exception:
        return None


# ------------------------------------------------------------------------------ CPU arm
def cpu_sample(n, sample_iters, threads, hierarchy=None, verbose=False):
    """Times `sample_iters` PCG+AMG iterations of the same workload with the oracle port:
    ParSpmmOp-tiled operator applies on `threads` OpenMP threads, serial CSR R/P and vector ops,
    unfused smooth/cycle with per-call allocation -- the reference's CPU path."""
    import oracle as O
    t0 = time.time()
    if hierarchy is None:
        a = O.gen_g7(n)
        nn = np.full((a.nrows, 1), 1.0 / np.sqrt(a.nrows))
        hierarchy = O.build_hierarchy(a, nn, (n, n, n))
    a = hierarchy.operators[0]
    mg = O.multigrid_from_hierarchy(hierarchy, "l1", nthreads=threads)
    par = O.ParSpmmOp(a, threads) if threads > 1 else None
    setup_s = time.time() - t0
    b = np.ones(a.nrows)

    def run():
        t = time.perf_counter()
        _, info = O.pcg(a, b, mg, rel_tol=REL_TOL, abs_tol=0.0, max_iters=sample_iters, par=par)
        return time.perf_counter() - t, info
    return run, setup_s, hierarchy


def scale_sample(seconds, sample_iters, full_iters):
    """A capped run does 1 + s preconditioner applies and s operator applies; a full solve of I
    iterations does I of each.  Per-iteration cost = t / (s + 1) (conservative for the CPU)."""
    return seconds / (sample_iters + 1) * full_iters


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm runs on rank 0 alone and
    # may use all the host threads it can (libgomp reads the variable when the oracle is loaded)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(cores)
    import oracle as O
    threads = O.num_threads()
    full = golden_iters(args.n) or 16
    s = args.sample_iters
    run, setup_s, _ = cpu_sample(args.n, s, threads)
    for _ in range(args.warmup):
        run()
    times = [run()[0] for _ in range(args.steps)]
    ms = scale_sample(float(np.mean(times)), s, full) * 1e3
    sample = (f"{s} of {full} PCG iterations (1+{s} V-cycles, {s} operator applies) of the same {args.n}^3 solve per step, "
              f"scaled x{full}/{s + 1}; hierarchy built by the oracle in {setup_s:.0f} s (untimed)")
    line = {"impl": "reference", "metric": "pcg_amg_solve_time", "value": ms, "unit": "ms", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": ms, "unit": "ms", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args):
    name = "7-point Poisson" if args.stencil == 7 else "27-point anisotropic diffusion (eps_z=1e-2)"
    return {"workload": f"3D {name} {args.n}^3 f64, PCG + smoothed-aggregation AMG V(1,1), L1-Jacobi smoother, "
                        f"rel_tol {REL_TOL:g}, b=1, zero guess (BASELINE configs[3]; configs[1] = 128^3 is a parity-test case)",
            "grid": [args.n] * 3, "rows": args.n ** 3, "rel_tol": REL_TOL, "l2": "inputs larger than L2 (no flush needed)",
            "partition": f"{args.gpus} z-slab(s)"}


# ------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import faer_amg_b200 as F
    from faer_amg_b200.distributed import Comm, DistMultigrid, level_row_splits

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: faer_amg_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = args.n
    ctx = F.Context.default(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)

    # ---- setup (untimed): operator + hierarchy on the device
    t0 = time.perf_counter()
    a = F.gallery.poisson7(ctx, n) if args.stencil == 7 else F.gallery.diffusion27(ctx, n)
    ctx.sync()
    t_gen = time.perf_counter() - t0
    rows = a.nrows
    nn = np.full((rows, 1), 1.0 / np.sqrt(rows))
    def build_hierarchy():
        t0 = time.perf_counter()
        gp = F.GeometricPartitioner((n, n, n), tuple(int(v) for v in args.block.split(",")))
        h = F.HierarchyConfig(1000, F.AggregationConfig(1, 1, gp)).build(F.SparseMatOp(a), nn)
        mg = F.MultigridConfig(smoother="l1").build(h)
        ctx.sync()
        return gp, h, mg, time.perf_counter() - t0
    # first build pays CUDA's lazy module loading of every setup kernel; the second is steady state
    gp, h, mg, t_setup_cold = build_hierarchy()
    del h, mg
    gp, h, mg, t_setup = build_hierarchy()
    params = F.CgParams(0.0, REL_TOL, 1000)

    if world > 1:
        comm = Comm.from_torch(ctx)
        splits = level_row_splits(gp.dims[: h.levels()], world)
        dmg = DistMultigrid(comm, mg, splits, replicate_below=args.replicate_below)
        nloc = dmg.nloc
    else:
        dmg, nloc = None, rows
    B = F.DeviceMat.from_host(ctx, np.ones(nloc))
    X = F.DeviceMat(ctx, nloc, 1)

    def solve_dev():
        return dmg.solve_dev(X, B, params) if dmg else F.conjugate_gradient_dev(X, mg, a, B, params)

    # pinned host buffers for the e2e (host-pointer ABI) path
    hb = torch.ones(nloc, dtype=torch.float64).pin_memory()
    hx = torch.zeros(nloc, dtype=torch.float64).pin_memory()
    hbn, hxn = hb.numpy(), hx.numpy()

    def solve_host():
        return dmg.solve(hxn, hbn, params) if dmg else F.conjugate_gradient(hxn, mg, a, hbn, params)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize; CUDA events on the launching stream; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count()
        e0.record(stream)
        infos = [fn() for _ in range(steps)]
        e1.record(stream)
        e1.synchronize()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, infos, ctx.launch_count() - l0

    for _ in range(max(args.warmup, 3)):
        solve_dev()
    with ClockSampler(local_rank) as clk:
        ms_dev, infos, launches = timed(solve_dev, args.steps)
    clocks = clk.summary()
    for _ in range(2):
        solve_host()
    ms_e2e, infos_h, _ = timed(solve_host, args.steps)
    iters = infos[0].iter_count

    line = None
    if rank == 0:
        peak, peak_src = peaks()
        # ---- kernel roofline on the fine level (N = 1 view of the dominant kernel)
        nnz = a.nnz
        bytes_spmv = 12.0 * nnz + 4.0 * (rows + 1) + 16.0 * rows
        bytes_resid = 12.0 * nnz + 4.0 * (rows + 1) + 24.0 * rows
        bytes_smooth = 12.0 * nnz + 4.0 * (rows + 1) + 32.0 * rows
        t_spmv = a.time_kernel(0, 50, 5)
        t_resid = a.time_kernel(1, 50, 5)
        t_smooth = a.time_kernel(2, 50, 5)
        gbs = lambda b, ms: b / (ms * 1e-3) / 1e9  # noqa: E731
        traffic = None
        try:  # DRAM bytes per launch from the committed ncu --set full capture of this exact operator
            if n == 256 and args.stencil == 7:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))["smooth"]["traffic_bytes_per_launch"]
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": "spmv_tma_kernel<TPR=1, EPI_SMOOTH> (fused x' = x + d.*(b - A x), fine level)",
                    "achieved": gbs(bytes_smooth, t_smooth), "peak": peak, "unit": "GB/s",
                    "frac": gbs(bytes_smooth, t_smooth) / peak, "frac_of_nominal_8TBs": gbs(bytes_smooth, t_smooth) / 8000.0,
                    "traffic": traffic, "peak_source": peak_src, "ms_per_launch": t_smooth, "algorithmic_bytes": bytes_smooth,
                    "spmv": {"GB/s": gbs(bytes_spmv, t_spmv), "ms": t_spmv, "frac": gbs(bytes_spmv, t_spmv) / peak},
                    "residual": {"GB/s": gbs(bytes_resid, t_resid), "ms": t_resid, "frac": gbs(bytes_resid, t_resid) / peak},
                    "cycle_algorithmic_bytes": mg.cycle_bytes(1)}
        line = {"metric": "pcg_amg_solve_time", "value": ms_dev, "unit": "ms", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": False, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args),
                "pcg_iterations": iters, "rel_residual": infos[0].rel_residual, "levels": h.levels(),
                "op_complexity": h.op_complexity(), "setup_ms": {"generate": t_gen * 1e3, "hierarchy_rap": t_setup * 1e3, "hierarchy_rap_first_call": t_setup_cold * 1e3},
                "mdof_per_s": rows / (ms_dev * 1e-3) / 1e6,
                "e2e": {"value": ms_e2e, "unit": "ms", "h2d_bytes_per_step": 8 * nloc, "d2h_bytes_per_step": 8 * nloc,
                        "api": "famg_pcg_solve (host pointers, pinned)" if not dmg else "famg_dist_pcg_solve (host pointers, pinned)",
                        "pcg_iterations": infos_h[0].iter_count},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
        # ---- CPU baseline on the host cores (bounded sample), rank 0, N = 1 only
        if world == 1 and not args.no_cpu:
            try:
                import oracle as O
                threads = O.num_threads()
                full = iters
                s = args.sample_iters
                oh = download_hierarchy(h)
                run, setup_s, _ = cpu_sample(n, s, threads, hierarchy=oh)
                run()
                t_cpu = float(np.mean([run()[0] for _ in range(2)]))
                ms_cpu = scale_sample(t_cpu, s, full) * 1e3
                line["cpu_baseline"] = {"value": ms_cpu, "unit": "ms", "cores": threads, "kind": "port",
                                        "sample": f"{s} of {full} PCG iterations of the same solve (hierarchy downloaded from the GPU "
                                                  f"build, bit-identical to the oracle's per tests), scaled x{full}/{s + 1}"}
            except Exception as exc:  # the baseline is a report, never a reason to lose the GPU line
                line["cpu_baseline"] = {"error": repr(exc)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def download_hierarchy(h):
    """GPU-built hierarchy -> oracle containers (input of the CPU baseline leg only)."""
    import oracle as O

    def conv(m):
        rp, ci, v = m.to_host()
        return O.Csr.from_arrays(m.nrows, m.ncols, rp.astype(np.int64), ci.astype(np.int64), v)
    oh = O.Hierarchy()
    oh.operators = [conv(op.mat_ref()) for op in h.operators()]
    oh.restrictions = [conv(r) for r in h.restrictions()]
    oh.interpolations = [conv(p) for p in h.interpolations()]
    return oh


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", "--n", dest="n", type=int, default=256,
                    help="grid points per dimension (use --grid under torchrun: its parser treats --n as an abbreviation)")
    ap.add_argument("--sample-iters", type=int, default=2, help="PCG iterations per CPU sample")
    ap.add_argument("--replicate-below", type=int, default=4096, help="rows per rank under which a level is replicated")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--stencil", type=int, default=7, choices=[7, 27], help="7-point Poisson (headline) or 27-point anisotropic diffusion")
    ap.add_argument("--block", default="2,2,2", help="geometric aggregate box")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
