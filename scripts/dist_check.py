"""Multi-rank correctness check of the row-partitioned path (run under torchrun, one rank per GPU):
distributed SpMV / V-cycle / PCG against the single-GPU result computed redundantly on each rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_amg_b200 as F  # noqa: E402
from faer_amg_b200.distributed import Comm, DistMultigrid, level_row_splits  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dims = tuple(int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (32, 32, 32)))
    rep = int(sys.argv[4]) if len(sys.argv) > 4 else 500
    ctx = F.Context.default(local)
    a = F.gallery.poisson7(ctx, *dims)
    n = a.nrows
    nn = np.full((n, 1), 1.0 / np.sqrt(n))
    gp = F.GeometricPartitioner(dims)
    h = F.HierarchyConfig(200, F.AggregationConfig(1, 1, gp)).build(F.SparseMatOp(a), nn)
    mg = F.MultigridConfig(smoother="l1").build(h)
    comm = Comm.from_torch(ctx)
    assert comm.allreduce_sum([1.0, rank])[0] == world
    splits = level_row_splits(gp.dims[: h.levels()], world)
    dmg = DistMultigrid(comm, mg, splits, replicate_below=rep)
    r0, r1 = dmg.row_begin, dmg.row_end
    rng = np.random.default_rng(0)
    xg = rng.standard_normal(n)
    ok = True
    # SpMV
    y = F.DeviceMat(ctx, r1 - r0, 1)
    dmg.spmv_dev(y, F.DeviceMat.from_host(ctx, xg[r0:r1]))
    ref = a.apply(xg).ravel()[r0:r1]
    ok &= bool(np.array_equal(y.to_host().ravel(), ref))
    print(f"[rank {rank}] rows {r0}:{r1} spmv bit-exact: {np.array_equal(y.to_host().ravel(), ref)}", flush=True)
    # V-cycle
    z = F.DeviceMat(ctx, r1 - r0, 1)
    dmg.apply_dev(z, F.DeviceMat.from_host(ctx, xg[r0:r1]))
    zref = mg.apply(xg).ravel()[r0:r1]
    err = np.max(np.abs(z.to_host().ravel() - zref)) / np.max(np.abs(zref))
    ok &= bool(err <= 1e-12)
    print(f"[rank {rank}] v-cycle rel err {err:.2e}", flush=True)
    # PCG
    b = np.ones(n)
    xs = np.zeros(n)
    i1 = F.conjugate_gradient(xs, mg, a, b, F.CgParams(0.0, 1e-10, 500))
    xl = np.zeros(r1 - r0)
    i2 = dmg.solve(xl, b[r0:r1], F.CgParams(0.0, 1e-10, 500))
    errx = np.linalg.norm(xl - xs[r0:r1]) / np.linalg.norm(xs)
    ok &= abs(i1.iter_count - i2.iter_count) <= 1 and errx <= 1e-9
    print(f"[rank {rank}] pcg iters single={i1.iter_count} dist={i2.iter_count} sol err {errx:.2e}", flush=True)
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_CHECK", "PASS" if t.item() == 1.0 else "FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
