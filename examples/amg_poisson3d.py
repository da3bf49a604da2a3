"""The call stack of examples/amg/main.rs on a synthetic 3-D Poisson problem:
SparseMatOp -> near-null -> HierarchyConfig::build (smoothed aggregation, GPU RAP) ->
MultigridConfig::build -> test_solver (PCG), printing iterations and operator complexity."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_amg_b200 as F  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    tol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-12          # examples/amg/main.rs:102-104
    ctx = F.Context.default(0)
    op = F.SparseMatOp(F.gallery.poisson7(ctx, n), block_size=1)
    rows = op.mat_ref().nrows
    near_null = np.full((rows, 1), 1.0 / np.sqrt(rows))              # constant vector, thin-Q normalised
    t0 = time.perf_counter()
    hierarchy = F.HierarchyConfig(1000, F.AggregationConfig(1, 1, F.GeometricPartitioner((n, n, n)))).build(op, near_null)
    mg = F.MultigridConfig(mu=1, smoothing_steps=1, smoother="l1").build(hierarchy)
    ctx.sync()
    print(hierarchy)
    print(f"setup {time.perf_counter() - t0:.2f} s")
    t0 = time.perf_counter()
    iters, rel, x = F.test_solver(op, mg, None, np.ones(rows), 1000, tol)
    print(f"cg_iters {iters}  rel_residual {rel:.2e}  op_complexity {hierarchy.op_complexity():.3f}  solve {time.perf_counter() - t0:.3f} s")


if __name__ == "__main__":
    main()
