"""examples/simple_geometric.rs of the reference, line for line, on the GPU path.

1-D Poisson -u'' = 1 on (0,1), finite differences, geometric multigrid (linear interpolation /
full weighting, Jacobi(0.66) smoothers, Cholesky on the coarsest level), V-cycle; compares PCG +
Jacobi, PCG + multigrid and stationary iteration + multigrid over refinements 2..10
(simple_geometric.rs:176-301).  BASELINE config #1.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_amg_b200 as F  # noqa: E402
from faer_amg_b200 import gallery  # noqa: E402


def main():
    ctx = F.Context.default(0)
    base_elements, max_iters, rel_tolerance = 10, 6000, 1e-8
    results = []
    print("=== 1D Geometric Multigrid Refinement Study ===")
    print(f"Target relative residual: {rel_tolerance:.0e}\n")
    for refinement in range(2, 11):
        n_elements = base_elements * 2 ** refinement
        n_dofs = n_elements - 1
        fine = gallery.poisson1d(ctx, n_elements)                       # make_finite_difference
        simple_pc = F.new_jacobi(fine, 0.66)
        mg = F.Multigrid(fine, F.new_jacobi(fine, 0.66))
        for level in range(1, refinement + 1):
            coarse_elements = base_elements * 2 ** (refinement - level)
            coarse_dofs = coarse_elements - 1
            r = gallery.restriction1d(ctx, coarse_dofs)                 # make_restriction
            p = gallery.interpolation1d(ctx, coarse_dofs)               # make_interpolation
            mat = gallery.poisson1d(ctx, coarse_elements)
            smoother = F.SparseCholeskySolve.new(mat) if level == refinement else F.new_jacobi(mat, 0.66)
            mg.add_level(mat, smoother, r, p)
        rhs = np.ones(n_dofs)
        params = F.CgParams(abs_tolerance=float(np.finfo(float).eps), rel_tolerance=rel_tolerance, max_iters=max_iters)
        dst = np.zeros(n_dofs)
        pcg_jacobi = F.conjugate_gradient(dst, simple_pc, fine, rhs, params).iter_count
        dst[:] = 0.0
        pcg_mg = F.conjugate_gradient(dst, mg, fine, rhs, params).iter_count
        dst[:] = 0.0
        stat_mg = F.stationary_solver(dst, rhs, fine, mg, max_iters, rel_tolerance)
        results.append((n_dofs, pcg_jacobi, pcg_mg, stat_mg))
        print(f"Refinement {refinement}: {n_elements} elements, {n_dofs} DOFs: {pcg_jacobi} / {pcg_mg} / {stat_mg}")
    print("\n=== RESULTS SUMMARY ===")
    print(f"{'DOFs':>8} {'PCG+Jacobi':>12} {'PCG+MG':>12} {'Stat+MG':>12}")
    print("-" * 64)
    for row in results:
        print("{:>8} {:>12} {:>12} {:>12}".format(*row))


if __name__ == "__main__":
    main()
