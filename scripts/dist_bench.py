"""Where the distributed cycle's time goes (run under torchrun, one rank per GPU): CUDA-event timings of the
distributed V-cycle, SpMV and PCG solve for every exchange mode (FAMG_OVERLAP = 0 | 1 | 2), for several depths of
the distributed part (replicate_below), next to the same cycle on ONE GPU working on a rank-sized grid (what the
local kernels alone cost).

    torchrun --nproc-per-node N scripts/dist_bench.py [GRID=256] [REPS=30]
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_amg_b200 as F  # noqa: E402
from faer_amg_b200.distributed import (Comm, DistGeometricPartitioner, DistHierarchy, DistMat, DistMultigrid,  # noqa: E402
                                       fine_plane_splits)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    dims = (n, n, n)
    ctx = F.Context.default(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    comm = Comm.from_torch(ctx)
    params = F.CgParams(0.0, 1e-8, 1000)

    def timed(fn, k):
        for _ in range(3):
            fn()
        ctx.sync(); torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(k):
            fn()
        e1.record(stream)
        e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / k], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    out = {"grid": n, "ranks": world, "rows_per_rank": n ** 3 // world, "runs": []}
    rs = fine_plane_splits(dims, world)
    plane = n * n
    quick = os.environ.get("DIST_BENCH_QUICK") == "1"
    for rep_below in ((4096,) if quick else (4096, 40000, 400000, 4000000)):
        slab = F.gallery.poisson7_slab(ctx, n, n, n, int(rs[rank]) // plane, int(rs[rank + 1]) // plane)
        a0 = DistMat.from_slabs(comm, [slab], n ** 3)
        nn = [np.full(int(rs[rank + 1] - rs[rank]), 1.0 / np.sqrt(n ** 3))]
        try:
            dh = DistHierarchy(comm, a0, nn, DistGeometricPartitioner(dims), coarsest_dim=1000, replicate_below=rep_below)
        except ValueError:
            break
        for mode, graph in ((("3", "1"), ("0", "1")) if quick else (("3", "1"), ("2", "1"), ("0", "1"), ("3", "0"))):
            if rep_below != 4096 and mode == "2":
                continue
            os.environ["FAMG_OVERLAP"] = mode
            os.environ["FAMG_DIST_GRAPH"] = graph
            dmg = DistMultigrid.from_hierarchy(comm, dh)
            nloc = dmg.nloc
            x, b, z = F.DeviceMat(ctx, nloc, 1), F.DeviceMat.from_host(ctx, np.ones(nloc)), F.DeviceMat(ctx, nloc, 1)
            t_cyc = timed(lambda: dmg.apply_dev(z, b), reps)
            t_spmv = timed(lambda: dmg.spmv_dev(z, b), reps) if graph == "1" else None
            info = dmg.solve_dev(x, b, params)
            t_solve = timed(lambda: dmg.solve_dev(x, b, params), 3)
            out["runs"].append({"replicate_below": rep_below, "distributed_levels": dh.n_dist, "overlap": int(mode), "graph": int(graph),
                                "cycle_us": 1e3 * t_cyc, "spmv_us": None if t_spmv is None else 1e3 * t_spmv, "solve_ms": t_solve,
                                "iters": info.iter_count, "us_per_iteration": 1e3 * t_solve / info.iter_count})
            if rank == 0:
                print(json.dumps(out["runs"][-1]), flush=True)
            del dmg
        del dh, a0
    # the local kernels alone: the same hierarchy depth on one GPU over a rank-sized grid
    d1 = (n, n, max(n // world, 2))
    a = F.gallery.poisson7(ctx, *d1)
    nn1 = np.full((a.nrows, 1), 1.0 / np.sqrt(a.nrows))
    h = F.HierarchyConfig(1000, F.AggregationConfig(1, 1, F.GeometricPartitioner(d1))).build(F.SparseMatOp(a), nn1)
    mg = F.MultigridConfig(smoother="l1").build(h)
    x, b, z = F.DeviceMat(ctx, a.nrows, 1), F.DeviceMat.from_host(ctx, np.ones(a.nrows)), F.DeviceMat(ctx, a.nrows, 1)
    t_cyc = timed(lambda: mg.apply_dev(z, b), reps)
    t_spmv = timed(lambda: a.apply_dev(z, b), reps)
    info = F.conjugate_gradient_dev(x, mg, a, b, params)
    t_solve = timed(lambda: F.conjugate_gradient_dev(x, mg, a, b, params), 3)
    out["single_gpu_rank_sized"] = {"grid": list(d1), "levels": h.levels(), "cycle_us": 1e3 * t_cyc, "spmv_us": 1e3 * t_spmv, "solve_ms": t_solve,
                                    "iters": info.iter_count, "us_per_iteration": 1e3 * t_solve / info.iter_count}
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
