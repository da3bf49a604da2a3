"""Multi-rank correctness check of the row-partitioned path (run under torchrun, one rank per GPU).

For both constructions of the distributed multigrid --
  "slabs":      the hierarchy built on row slabs (DistHierarchy: every rank computes only its rows of P, R, A_c)
  "replicated": slabs cut from a hierarchy every rank built in full (round-1 path, kept for the C ABI)
-- the distributed SpMV, V-cycle and PCG are compared with the single-GPU result computed redundantly on each rank:
SpMV and V-cycle bit for bit, PCG iteration counts within +-1, and (slabs) every P / R / A_c slab bit for bit
against the rows of the undistributed build.

    torchrun --nproc-per-node N scripts/dist_check.py NX NY NZ REPLICATE_BELOW [STENCIL]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_amg_b200 as F  # noqa: E402
from faer_amg_b200.distributed import (Comm, DistGeometricPartitioner, DistHierarchy, DistMat, DistMultigrid,  # noqa: E402
                                       fine_plane_splits, level_row_splits)


def rows_of(m, r0, r1):
    rp, ci, v = m.to_host()
    rp = rp.astype(np.int64)
    return rp[r0:r1 + 1] - rp[r0], ci[rp[r0]:rp[r1]].astype(np.int64), v[rp[r0]:rp[r1]]


def same_rows(slab, glob, r0, r1):
    rp, ci, v = slab.to_host()
    grp, gci, gv = rows_of(glob, r0, r1)
    return (np.array_equal(rp.astype(np.int64), grp) and np.array_equal(ci.astype(np.int64), gci) and np.array_equal(v, gv))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dims = tuple(int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (32, 32, 32)))
    rep = int(sys.argv[4]) if len(sys.argv) > 4 else 500
    stencil = int(sys.argv[5]) if len(sys.argv) > 5 else 7
    coarsest = 200
    ctx = F.Context.default(local)
    a = F.gallery.poisson7(ctx, *dims) if stencil == 7 else F.gallery.diffusion27(ctx, *dims)
    n = a.nrows
    nn = np.full((n, 1), 1.0 / np.sqrt(n))
    gp = F.GeometricPartitioner(dims)
    h = F.HierarchyConfig(coarsest, F.AggregationConfig(1, 1, gp)).build(F.SparseMatOp(a), nn)
    mg = F.MultigridConfig(smoother="l1").build(h)
    comm = Comm.from_torch(ctx)
    assert comm.allreduce_sum([1.0, rank])[0] == world
    rng = np.random.default_rng(0)
    xg = rng.standard_normal(n)
    b = np.ones(n)
    xs = np.zeros(n)
    i1 = F.conjugate_gradient(xs, mg, a, b, F.CgParams(0.0, 1e-10, 500))
    ref_spmv = a.apply(xg).ravel()
    ref_cycle = mg.apply(xg).ravel()
    ok = True

    def check(tag, dmg):
        nonlocal ok
        r0, r1 = dmg.row_begin, dmg.row_end
        y = F.DeviceMat(ctx, r1 - r0, 1)
        dmg.spmv_dev(y, F.DeviceMat.from_host(ctx, xg[r0:r1]))
        e_spmv = bool(np.array_equal(y.to_host().ravel(), ref_spmv[r0:r1]))
        z = F.DeviceMat(ctx, r1 - r0, 1)
        dmg.apply_dev(z, F.DeviceMat.from_host(ctx, xg[r0:r1]))
        zz = z.to_host().ravel()
        e_cyc = bool(np.array_equal(zz, ref_cycle[r0:r1]))
        err = np.max(np.abs(zz - ref_cycle[r0:r1])) / np.max(np.abs(ref_cycle))
        xl = np.zeros(r1 - r0)
        i2 = dmg.solve(xl, b[r0:r1], F.CgParams(0.0, 1e-10, 500))
        errx = np.linalg.norm(xl - xs[r0:r1]) / np.linalg.norm(xs)
        good = e_spmv and err <= 1e-12 and abs(i1.iter_count - i2.iter_count) <= 1 and errx <= 1e-9
        ok &= good
        print(f"[rank {rank}] {tag}: rows {r0}:{r1} spmv bit-exact {e_spmv}, v-cycle bit-exact {e_cyc} (rel err {err:.2e}), "
              f"pcg iters single={i1.iter_count} dist={i2.iter_count}, sol err {errx:.2e} -> {'ok' if good else 'FAIL'}", flush=True)

    # ---- hierarchy built on row slabs
    rs = fine_plane_splits(dims, world)
    plane = dims[0] * dims[1]
    gen = F.gallery.poisson7_slab if stencil == 7 else F.gallery.diffusion27_slab
    slab = gen(ctx, *dims, int(rs[rank]) // plane, int(rs[rank + 1]) // plane)
    a0 = DistMat.from_slabs(comm, [slab], n)
    dh = DistHierarchy(comm, a0, [nn[rs[rank]:rs[rank + 1], 0].copy()], DistGeometricPartitioner(dims), coarsest_dim=coarsest,
                       replicate_below=rep)
    slabs_ok = dh.levels() == h.levels()
    for l in range(dh.n_dist):
        rsl, csl = dh.A[l].row_split(), dh.P[l].info()[3]
        slabs_ok &= same_rows(dh.A[l].local(0, True), h.get_mat_ref(l), int(rsl[rank]), int(rsl[rank + 1]))
        slabs_ok &= same_rows(dh.P[l].local(0, True), h.get_interpolation(l), int(rsl[rank]), int(rsl[rank + 1]))
        slabs_ok &= same_rows(dh.R[l].local(0, True), h.get_restriction(l), int(csl[rank]), int(csl[rank + 1]))
    for t in range(dh.tail.levels()):
        w = h.get_mat_ref(dh.n_dist + t)
        slabs_ok &= same_rows(dh.tail.get_mat_ref(t), w, 0, w.nrows)
    ok &= bool(slabs_ok)
    print(f"[rank {rank}] slabs: {dh.n_dist} distributed + {dh.tail.levels()} replicated levels, P/R/A_c slabs bit-identical: {slabs_ok}", flush=True)
    check("slabs", DistMultigrid.from_hierarchy(comm, dh))
    # ---- slabs cut from the replicated hierarchy
    splits = level_row_splits(gp.dims[: h.levels()], world)
    check("replicated", DistMultigrid(comm, mg, splits, replicate_below=rep))

    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_CHECK", "PASS" if t.item() == 1.0 else "FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
