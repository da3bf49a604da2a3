import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import faer_amg_b200 as F
ctx = F.Context.default(0)
n = 128
a = F.gallery.poisson7(ctx, n)
rows = a.nrows
nn = np.full((rows, 1), 1.0 / np.sqrt(rows))
t=time.perf_counter()
h = F.HierarchyConfig(1000, F.AggregationConfig(1, 1, F.GeometricPartitioner((n, n, n)))).build(F.SparseMatOp(a), nn)
ctx.sync(); print("hierarchy", time.perf_counter()-t)
mg = F.MultigridConfig(smoother="l1").build(h)
params = F.CgParams(0.0, 1e-8, 1000)
hb = torch.ones(rows, dtype=torch.float64).pin_memory().numpy(); hx = torch.zeros(rows, dtype=torch.float64).pin_memory().numpy()
B = F.DeviceMat.from_host(ctx, hb); X = F.DeviceMat(ctx, rows, 1)
for i in range(3):
    t=time.perf_counter(); F.conjugate_gradient_dev(X, mg, a, B, params); print("dev solve", time.perf_counter()-t)
for i in range(3):
    t=time.perf_counter(); F.conjugate_gradient(hx, mg, a, hb, params); print("host solve", time.perf_counter()-t)
for i in range(3):
    t=time.perf_counter(); B.upload(hb); print("upload", time.perf_counter()-t)
    t=time.perf_counter(); y = X.to_host(); print("download(to new numpy)", time.perf_counter()-t)
    t=time.perf_counter(); v = F.DeviceMat(ctx, rows, 1); print("vec create", time.perf_counter()-t)
    t=time.perf_counter(); del v; print("vec destroy", time.perf_counter()-t)
# per-level SpGEMM timing
from faer_amg_b200.interpolation import tentative_prolongator, smooth_interpolation
part,_ = F.geometric_partition((n,n,n))
t=time.perf_counter(); p0, cnn = tentative_prolongator(ctx, rows, part, nn, 1); ctx.sync(); print("tentative_p", time.perf_counter()-t)
t=time.perf_counter(); p = smooth_interpolation(a, p0); ctx.sync(); print("smooth_interp", time.perf_counter()-t)
t=time.perf_counter(); r = p.transpose(); ctx.sync(); print("transpose", time.perf_counter()-t)
t=time.perf_counter(); ap = a @ p; ctx.sync(); print("A*P", time.perf_counter()-t)
t=time.perf_counter(); ac = r @ ap; ctx.sync(); print("R*AP", time.perf_counter()-t, ac.nnz)
t=time.perf_counter(); s = F.new_l1(ac); ctx.sync(); print("l1", time.perf_counter()-t)
t=time.perf_counter(); part2,_ = F.geometric_partition((n//2,)*3); print("partition py", time.perf_counter()-t)
