"""How expensive is device memory on this box?  Times first-touch growth of the stream-ordered pool (famg_ctx_reserve),
re-allocation out of the retained pool, and a hierarchy build before / after pre-sizing the pool."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import faer_amg_b200 as F
ctx = F.Context.default(0)
def t(fn):
    ctx.sync(); t0 = time.perf_counter(); fn(); ctx.sync(); return 1e3 * (time.perf_counter() - t0)
for gb in (1, 1, 4, 4, 16, 16):
    print(f"reserve {gb:3d} GB: {t(lambda: ctx.reserve(gb << 30)):9.2f} ms", flush=True)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
a = F.gallery.poisson7(ctx, n)
nn = np.full((a.nrows, 1), 1.0 / np.sqrt(a.nrows))
for rep in range(4):
    ms = t(lambda: F.HierarchyConfig(1000, F.AggregationConfig(1, 1, F.GeometricPartitioner((n, n, n)))).build(F.SparseMatOp(a), nn))
    print(f"hierarchy build {rep}: {ms:9.2f} ms", flush=True)
