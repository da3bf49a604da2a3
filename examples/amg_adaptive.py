"""The algebraic half of examples/amg/main.rs (:325-474) on a synthetic 27-point anisotropic operator:
find_near_null (device-resident smooth_vector: fused error-propagator SpMM + CholeskyQR2, block
smoother over algebraic aggregates) -> constant + smooth basis, thin Q, create_weights ->
HierarchyConfig::build with the host partitioner (PartitionerConfig) and candidate_dimension > 1
(block_jacobi prolongator smoothing on the coarse levels) -> MultigridConfig::build -> test_solver;
with --composite N the AdaptiveConfig::build loop (:325-346) instead.
usage: amg_adaptive.py [N=24] [near_null_dim=4] [--composite 2]
The defaults are the configuration exercised by tests/test_gpu_composite.py (near-null width 4,
candidate_dimension 2, algebraic aggregates of ~8 nodes)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_amg_b200 as F  # noqa: E402
from faer_amg_b200.hierarchy import thin_q  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n = int(args[0]) if len(args) > 0 else 24
    nn_dim = int(args[1]) if len(args) > 1 else 4                      # --coarsening-near-null-dim
    composite = int(sys.argv[sys.argv.index("--composite") + 1]) if "--composite" in sys.argv else 0
    smoothing_iters, block_smoother_size, tol = 10, 16.0, 1e-12        # examples/amg/main.rs:53-58, 102-108
    ctx = F.Context.default(0)
    op = F.SparseMatOp(F.gallery.diffusion27(ctx, n, n, n, 1.0, 1e-2), block_size=1)
    rows = op.mat_ref().nrows
    hierarchy_config = F.HierarchyConfig(1000, F.AggregationConfig(1, min(nn_dim, 2), F.PartitionerConfig(8.0, 1.0, 50)))
    multigrid_config = F.MultigridConfig(mu=1, smoothing_steps=1, smoother="l1")
    b = np.ones(rows)
    t0 = time.perf_counter()
    if composite:
        pc = F.AdaptiveConfig(hierarchy_config, multigrid_config, max_components=composite, test_iters=smoothing_iters,
                              coarsening_near_null_dim=nn_dim, smoothing_block_size=block_smoother_size, seed=42).build(op)
        print(f"composite of {len(pc.components())} multigrids, setup {time.perf_counter() - t0:.2f} s")
        complexity = float("nan")
    else:
        nn = F.find_near_null(op, smoothing_iters, nn_dim - 1, block_smoother_size, seed=42)          # :348-353
        nn_with_constant = np.ones((rows, nn_dim), order="F")
        nn_with_constant[:, 1:] = nn                                                                   # :355-359
        basis = thin_q(nn_with_constant)
        weights = F.create_weights(basis, op)                                                          # :377
        hierarchy = hierarchy_config.build(op, basis, weights)
        print(hierarchy)
        pc = multigrid_config.build(hierarchy)
        complexity = hierarchy.op_complexity()
        print(f"setup {time.perf_counter() - t0:.2f} s")
    t0 = time.perf_counter()
    iters, rel, _ = F.test_solver(op, pc, None, b, 1000, tol)
    print(f"cg_iters {iters}  rel_residual {rel:.2e}  op_complexity {complexity:.3f}  solve {time.perf_counter() - t0:.3f} s")


if __name__ == "__main__":
    main()
