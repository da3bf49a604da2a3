"""Phase timings of the distributed hierarchy build (run under torchrun with FAMG_SETUP_TRACE=1): second build, rank 0."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_amg_b200 as F  # noqa: E402
from faer_amg_b200.distributed import Comm, DistGeometricPartitioner, DistHierarchy, DistMat, DistMultigrid, fine_plane_splits  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = [int(v) for v in sys.argv[1:4]] if len(sys.argv) > 3 else [int(sys.argv[1]) if len(sys.argv) > 1 else 256] * 3
dims = tuple(g)
n = dims[0]
ctx = F.Context.default(local)
comm = Comm.from_torch(ctx)
rs = fine_plane_splits(dims, world)
plane = dims[0] * dims[1]
ntot = dims[0] * dims[1] * dims[2]
for rep in range(3):
    dist.barrier()
    if rank == 0:
        print(f"[setup] ---- build {rep}", file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    slab = F.gallery.poisson7_slab(ctx, dims[0], dims[1], dims[2], int(rs[rank]) // plane, int(rs[rank + 1]) // plane)
    a0 = DistMat.from_slabs(comm, [slab], ntot)
    nn = [np.full(int(rs[rank + 1] - rs[rank]), 1.0 / np.sqrt(ntot))]
    ctx.sync(); t1 = time.perf_counter()
    dh = DistHierarchy(comm, a0, nn, DistGeometricPartitioner(dims), coarsest_dim=1000, replicate_below=4096)
    ctx.sync(); t2 = time.perf_counter()
    dmg = DistMultigrid.from_hierarchy(comm, dh)
    ctx.sync(); t3 = time.perf_counter()
    if rank == 0:
        print(f"[setup] build {rep}: slab {1e3*(t1-t0):.1f} ms, hierarchy {1e3*(t2-t1):.1f} ms, multigrid + peer-memory setup {1e3*(t3-t2):.1f} ms",
              file=sys.stderr, flush=True)
    del dmg, dh, a0
dist.barrier()
dist.destroy_process_group()
