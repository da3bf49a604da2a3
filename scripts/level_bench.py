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
# transfer operators: f_c = R w (EPI_SPMV) and v += P v_c (EPI_ADD)
def timeit(fn, reps=50):
    for _ in range(5): fn()
    ctx.sync(); t = time.perf_counter()
    for _ in range(reps): fn()
    ctx.sync(); return (time.perf_counter() - t) / reps * 1e6
for lvl in range(h.levels() - 1):
    P, R = h.get_interpolation(lvl), h.get_restriction(lvl)
    xf, xc = F.DeviceMat(ctx, P.nrows, 1), F.DeviceMat(ctx, P.ncols, 1)
    xf.fill(1.0); xc.fill(1.0)
    tp = timeit(lambda: P.apply_add_dev(xf, xc)); tr = timeit(lambda: R.apply_dev(xc, xf))
    bp = 12.0 * P.nnz + 4.0 * (P.nrows + 1) + 8.0 * P.ncols + 16.0 * P.nrows
    br = 12.0 * R.nnz + 4.0 * (R.nrows + 1) + 8.0 * R.ncols + 8.0 * R.nrows
    print(f"transfer {lvl}->{lvl+1}: P {P.nrows}x{P.ncols} {P.nnz/P.nrows:.1f}/row tpr {P.plan()['threads_per_row']} {tp:7.1f} us {bp/tp/1e3:6.0f} GB/s | "
          f"R {R.nnz/R.nrows:.1f}/row tpr {R.plan()['threads_per_row']} {tr:7.1f} us {br/tr/1e3:6.0f} GB/s", flush=True)
Z, Rh = F.DeviceMat(ctx, a.nrows, 1), F.DeviceMat(ctx, a.nrows, 1); Rh.fill(1.0)
mg = F.MultigridConfig(smoother="l1").build(h)
print(f"v-cycle {timeit(lambda: mg.apply_dev(Z, Rh), 20):.0f} us for {mg.cycle_bytes(1)/1e9:.2f} GB algorithmic")
