"""Near-null search loop (adaptivity.rs:351-354) on the device: per-step time of the fused E apply
(SpMM of width k + x - d.*(Ax)) and of the CholeskyQR2 thin Q, against their algorithmic bytes
(tsqr.cu header: E apply 12 nnz + 4(n+1) + 8n + 16nk; thin Q 48nk), and the same step done the
reference's way on the host cores (oracle SpMM + Householder QR) for a bounded sample.
usage: nearnull_bench.py [N=128] [k=16,64]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import faer_amg_b200 as F
from faer_amg_b200._ffi import call
from faer_amg_b200.hierarchy import thin_q_dev

ctx = F.Context.default(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ks = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [16, 64]
peak = 6550.4
a = F.gallery.poisson7(ctx, N)
n, nnz = a.nrows, a.nnz
l1 = F.new_l1(a)
E = F.ErrorPropogator(F.SparseMatOp(a), l1)


def timeit(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    ctx.sync(); t = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.sync()
    return (time.perf_counter() - t) / reps * 1e3


for k in ks:
    x0 = np.random.default_rng(0).standard_normal((n, k))
    X, Y = F.DeviceMat.from_host(ctx, x0), F.DeviceMat(ctx, n, k)
    thin_q_dev(X)
    t_e = timeit(lambda: E.apply_dev(Y, X))
    t_q = timeit(lambda: thin_q_dev(X))
    b_e = 12.0 * nnz + 4.0 * (n + 1) + 8.0 * n + 16.0 * n * k
    b_q = 48.0 * n * k
    cfs = np.zeros(k)
    Z = F.DeviceMat.from_host(ctx, x0)
    ctx.sync(); t = time.perf_counter()
    call("famg_smooth_vector_dev", a._h, l1._h, 10, Z._h, cfs.ctypes.data_as(F._ffi.f64p))
    ctx.sync(); t_sv = (time.perf_counter() - t) * 1e3
    print(f"G7({N}) k={k}: E apply {t_e:.3f} ms = {b_e / t_e / 1e6:.0f} GB/s ({b_e / t_e / 1e6 / peak:.0%}) | thin Q {t_q:.3f} ms = "
          f"{b_q / t_q / 1e6:.0f} GB/s ({b_q / t_q / 1e6 / peak:.0%}) | smooth_vector(10 it) {t_sv:.1f} ms, cf[0]={cfs[0]:.3f}", flush=True)

if os.environ.get("NEARNULL_CPU", "1") == "1":
    import oracle as O
    o = O.gen_g7(min(N, 64))
    d = O.new_l1(o).reshape(-1, 1)
    for k in ks:
        x = O.thin_q(np.random.default_rng(0).standard_normal((o.nrows, k)))
        t = time.perf_counter(); x = O.error_propagator(o, lambda r: d * r, x); te = time.perf_counter() - t
        t = time.perf_counter(); x = O.thin_q(x); tq = time.perf_counter() - t
        print(f"CPU port G7({min(N, 64)}) k={k} ({O.num_threads()} threads): E apply {te * 1e3:.1f} ms, Householder thin Q {tq * 1e3:.1f} ms", flush=True)
