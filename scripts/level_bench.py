"""Per-level kernel timing of a G7(n) hierarchy: SpMV / residual / smoother sweep on every level's
operator (CUDA events, 100 launches), plus setup wall time (second, warm build)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import faer_amg_b200 as F
ctx = F.Context.default(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
peak = 6550.4
a = F.gallery.poisson7(ctx, n)
nn = np.full((a.nrows, 1), 1.0 / np.sqrt(a.nrows))
for rep in range(2):
    ctx.sync(); t0 = time.perf_counter()
    h = F.HierarchyConfig(1000, F.AggregationConfig(1, 1, F.GeometricPartitioner((n, n, n)))).build(F.SparseMatOp(a), nn)
    ctx.sync(); print(f"hierarchy build {rep}: {1e3*(time.perf_counter()-t0):.1f} ms", flush=True)
print(os.environ.get("FAMG_LIB", "default lib"))
for lvl in range(h.levels() - 1):
    m = h.get_mat_ref(lvl)
    rows, nnz = m.nrows, m.nnz
    by = {0: 12.0 * nnz + 4.0 * (rows + 1) + 16.0 * rows, 1: 12.0 * nnz + 4.0 * (rows + 1) + 24.0 * rows, 2: 12.0 * nnz + 4.0 * (rows + 1) + 32.0 * rows}
    out = []
    for which, name in ((0, "spmv"), (1, "resid"), (2, "smooth")):
        ms = m.time_kernel(which, 100, 10)
        out.append(f"{name} {ms*1e3:7.1f} us {by[which]/(ms*1e-3)/1e9:6.0f} GB/s")
    print(f"level {lvl}: rows {rows} nnz/row {nnz/rows:.1f} tpr {m.plan()['threads_per_row']} " + " | ".join(out), flush=True)
