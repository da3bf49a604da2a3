"""Hierarchy::coarsen + PCG on several GPUs, every rank holding only its row slab (run under torchrun):

    torchrun --nproc-per-node 4 --master-addr 127.0.0.1 examples/dist_hierarchy.py 128

Each rank generates its z-slab of the 7-point operator, the hierarchy is built on the slabs (P, R, A_c rows computed where they
live, off-rank rows of P / AP exchanged), small coarse levels are gathered and run replicated, and the PCG + AMG solve exchanges
halos as peer-memory stores.  The same script on one GPU can host all the ranks of the *construction* in one process:

    python examples/dist_hierarchy.py 32 --virtual-ranks 4
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_amg_b200 as F  # noqa: E402
from faer_amg_b200.distributed import (Comm, DistGeometricPartitioner, DistHierarchy, DistMat, DistMultigrid,  # noqa: E402
                                       fine_plane_splits)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    dims = (n, n, n)
    plane, total = n * n, n ** 3
    if "--virtual-ranks" in sys.argv:  # construction only, one process
        nv = int(sys.argv[sys.argv.index("--virtual-ranks") + 1])
        ctx = F.Context.default(0)
        comm = Comm.sim(ctx, nv)
        rs = fine_plane_splits(dims, nv)
        slabs = [F.gallery.poisson7_slab(ctx, n, n, n, int(rs[r]) // plane, int(rs[r + 1]) // plane) for r in range(nv)]
        nn = [np.full(int(rs[r + 1] - rs[r]), 1.0 / np.sqrt(total)) for r in range(nv)]
        dh = DistHierarchy(comm, DistMat.from_slabs(comm, slabs, total), nn, DistGeometricPartitioner(dims), coarsest_dim=200,
                           replicate_below=max(256, total // nv // 64))
        print(f"{nv} virtual ranks: {dh.n_dist} distributed + {dh.tail.levels()} replicated levels; rows per level:",
              [m.nrows for m in dh.A] + [dh.tail.get_mat_ref(t).nrows for t in range(dh.tail.levels())])
        return
    import torch
    import torch.distributed as dist
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = F.Context.default(local)
    comm = Comm.from_torch(ctx)
    rs = fine_plane_splits(dims, world)
    slab = F.gallery.poisson7_slab(ctx, n, n, n, int(rs[rank]) // plane, int(rs[rank + 1]) // plane)
    nloc = int(rs[rank + 1] - rs[rank])
    dh = DistHierarchy(comm, DistMat.from_slabs(comm, [slab], total), [np.full(nloc, 1.0 / np.sqrt(total))], DistGeometricPartitioner(dims))
    dmg = DistMultigrid.from_hierarchy(comm, dh, smoother="l1")
    x = np.zeros(nloc)
    info = dmg.solve(x, np.ones(nloc), F.CgParams(0.0, 1e-8, 1000))
    if rank == 0:
        print(f"{world} ranks, {total} rows: {dh.n_dist} distributed + {dh.tail.levels()} replicated levels, "
              f"{info.iter_count} PCG iterations, rel. residual {info.rel_residual:.2e}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
