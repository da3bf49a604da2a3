"""Timing of the distributed pieces vs their single-GPU counterparts (run under torchrun)."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_amg_b200 as F
from faer_amg_b200.distributed import Comm, DistMultigrid, level_row_splits
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
ctx = F.Context.default(local)
a = F.gallery.poisson7(ctx, n)
rows = a.nrows
nn = np.full((rows, 1), 1.0 / np.sqrt(rows))
gp = F.GeometricPartitioner((n, n, n))
h = F.HierarchyConfig(1000, F.AggregationConfig(1, 1, gp)).build(F.SparseMatOp(a), nn)
mg = F.MultigridConfig(smoother="l1").build(h)
comm = Comm.from_torch(ctx)
dmg = DistMultigrid(comm, mg, level_row_splits(gp.dims[: h.levels()], world), replicate_below=rep)
nloc = dmg.nloc
x, y = F.DeviceMat.from_host(ctx, np.ones(nloc)), F.DeviceMat(ctx, nloc, 1)
X, Y = F.DeviceMat.from_host(ctx, np.ones(rows)), F.DeviceMat(ctx, rows, 1)
def timeit(fn, reps=50):
    for _ in range(5): fn()
    ctx.sync(); dist.barrier(); t = time.perf_counter()
    for _ in range(reps): fn()
    ctx.sync(); t = (time.perf_counter() - t) / reps
    tt = torch.tensor([t], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return tt.item() * 1e6
t_spmv_d = timeit(lambda: dmg.spmv_dev(y, x))
t_spmv_1 = timeit(lambda: a.apply_dev(Y, X))
t_cyc_d = timeit(lambda: dmg.apply_dev(y, x), 20)
t_cyc_1 = timeit(lambda: mg.apply_dev(Y, X), 20)
params = F.CgParams(0.0, 1e-8, 1000)
xs = F.DeviceMat(ctx, nloc, 1)
t_solve = timeit(lambda: dmg.solve_dev(xs, x, params), 3)
if rank == 0:
    print(f"world {world} n {n} replicate_below {rep} graph {os.environ.get('FAMG_DIST_GRAPH','1')}: "
          f"spmv dist {t_spmv_d:.0f} us (1-GPU full {t_spmv_1:.0f} us) | v-cycle dist {t_cyc_d:.0f} us (1-GPU {t_cyc_1:.0f} us) | solve {t_solve/1e3:.2f} ms", flush=True)
dist.barrier(); dist.destroy_process_group()
