import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import faer_amg_b200 as F
from faer_amg_b200.interpolation import tentative_prolongator, smooth_interpolation
from faer_amg_b200.hierarchy import thin_q
ctx = F.Context.default(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
a = F.gallery.poisson7(ctx, n)
rows = a.nrows
nn = np.full((rows, 1), 1.0 / np.sqrt(rows))
def T(label, fn):
    ctx.sync(); t = time.perf_counter(); r = fn(); ctx.sync(); print(f"  {label:28s} {1e3*(time.perf_counter()-t):9.2f} ms", flush=True); return r
for rep in range(2):
    print("build", rep)
    t0 = time.perf_counter()
    dims = (n, n, n); mat = a; near = nn
    lvl = 0
    while mat.nrows > 1000:
        print(f" level {lvl}: rows {mat.nrows} nnz {mat.nnz}")
        part, cd = T("partition (numpy)", lambda: F.geometric_partition(dims))
        p0, cnn = T("tentative_p (host svd+upload)", lambda: tentative_prolongator(ctx, mat.nrows, part, near, 1))
        p = T("smooth_interpolation", lambda: smooth_interpolation(mat, p0))
        r = T("transpose", lambda: p.transpose())
        ap = T("A*P", lambda: mat @ p)
        ac = T("R*(AP)", lambda: r @ ap)
        l1 = T("new_l1", lambda: F.new_l1(ac))
        dv = T("upload nn", lambda: F.DeviceMat.from_host(ctx, cnn))
        T("stationary it", lambda: F.StationaryIteration(ac, l1, 3).apply_in_place_dev(dv))
        near = T("download+thin_q", lambda: thin_q(dv.to_host()))
        mat = ac; dims = cd; lvl += 1
    T("cholesky coarse", lambda: F.SparseCholeskySolve.new(mat))
    print(" total", time.perf_counter() - t0)
