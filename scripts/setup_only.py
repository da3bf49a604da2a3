import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import faer_amg_b200 as F
ctx = F.Context.default(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
a = F.gallery.poisson7(ctx, n)
nn = np.full((a.nrows, 1), 1.0 / np.sqrt(a.nrows))
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
for r in range(reps):
    ctx.sync(); t = time.perf_counter()
    h = F.HierarchyConfig(1000, F.AggregationConfig(1, 1, F.GeometricPartitioner((n, n, n)))).build(F.SparseMatOp(a), nn)
    ctx.sync(); print("hierarchy", time.perf_counter() - t, flush=True)
