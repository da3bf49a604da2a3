"""Times the three fused fine-level kernels (SpMV, residual, smoother sweep) with CUDA events for
both kernel variants and several sizes; prints GB/s against the measured HBM peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_amg_b200 as F
ctx = F.Context.default(0)
peak = 6550.4
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
sizes = [int(v) for v in sys.argv[1:]] or [128, 256]
variants = [int(v) for v in os.environ.get("VARIANTS", "1,2").split(",")]
for n in sizes:
    for stencil in ("g7", "g27") if n <= 192 else ("g7",):
        a = F.gallery.poisson7(ctx, n) if stencil == "g7" else F.gallery.diffusion27(ctx, n)
        rows, nnz = a.nrows, a.nnz
        by = {0: 12.0 * nnz + 4.0 * (rows + 1) + 16.0 * rows, 1: 12.0 * nnz + 4.0 * (rows + 1) + 24.0 * rows,
              2: 12.0 * nnz + 4.0 * (rows + 1) + 32.0 * rows}
        for v in variants:
            ctx.set_option("spmv_variant", v)
            out = []
            for which, name in ((0, "spmv"), (1, "resid"), (2, "smooth")):
                ms = a.time_kernel(which, 100, 10)
                gbs = by[which] / (ms * 1e-3) / 1e9
                out.append(f"{name} {ms*1e3:7.1f} us {gbs:6.0f} GB/s ({gbs/peak*100:4.1f}% of measured, {gbs/80:4.1f}% of 8TB/s)")
            print(f"{stencil} n={n} tpr={a.plan()['threads_per_row']} variant={v}: " + " | ".join(out), flush=True)
        del a
