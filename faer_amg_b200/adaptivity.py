"""The pieces of ``src/adaptivity.rs`` that sit on the hot path: ``ErrorPropogator`` (sic), the
operator ``E = I - M^-1 A`` whose ``apply`` (adaptivity.rs:191-198) is the inner loop of the
near-null search (``:351-354``) and of compatible relaxation (``interpolation/mod.rs:623-625``),
and the search loop itself, ``smooth_vector`` / ``find_near_null`` / ``create_weights``
(``:264-390, 434-443``), kept resident in HBM: SpMM of width k with the fused
``x - d .* (A x)`` epilogue plus a device CholeskyQR2 per step (``csrc/tsqr.cu``).
``AdaptiveConfig`` (``:28-165``) is the thin host loop over these pieces and ``Composite``
(``preconditioners/composite.py``), SURVEY 8f-4."""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np

from ._ffi import call
from .core import DeviceMat, ParSpmmOp, SparseMatOp, SparseRowMat, _f
from .preconditioners.smoothers import Smoother, new_l1


def _mat_of(op) -> SparseRowMat:
    return op.mat_ref() if isinstance(op, SparseMatOp) else op.mat if isinstance(op, ParSpmmOp) else op


class ErrorPropogator:
    """``out = x - M^-1 (A x)`` on n x k blocks, device-resident.  With a smoother of this package
    (``Diag``, ``BlockSmoother`` ...) the apply is one C-ABI call -- a single fused SpMM launch for
    ``Diag``; any other preconditioner object with ``apply_dev`` (e.g. ``Multigrid``) takes the
    three-step path."""

    def __init__(self, op, pc):
        self.op = _mat_of(op)
        self.pc = pc

    def nrows(self) -> int:
        return self.op.nrows

    ncols = nrows

    def apply_dev(self, out: DeviceMat, x: DeviceMat, work: Optional[DeviceMat] = None):
        if isinstance(self.pc, Smoother):
            call("famg_error_propagator_dev", self.op._h, self.pc._h, out._h, x._h)
            return
        if work is None:
            work = DeviceMat(self.op.ctx, x.nrows, x.ncols)
        self.op.apply_dev(work, x)          # work = A x
        self.pc.apply_dev(out, work)        # out = M^-1 work
        out.axpby(1.0, x, -1.0)             # out = x - out

    def apply(self, x) -> np.ndarray:
        ctx = self.op.ctx
        X = DeviceMat.from_host(ctx, x)
        out = DeviceMat(ctx, X.nrows, X.ncols)
        self.apply_dev(out, X)
        return out.to_host()

    conj_apply = apply


def smooth_vector_dev(mat, pc, iterations: int, x: DeviceMat, report: bool = False) -> List[float]:
    """adaptivity.rs:307-390 on a device-resident block ``x`` (n x k, k <= 64), in place; returns the
    per-column convergence factors ``||E w||_A / ||w||_A`` (``:365-384``).  ``pc``: a smoother, a
    ``Multigrid`` or a ``Composite`` (the adaptive driver passes the latter, ``:104-114``)."""
    from .preconditioners.composite import pc_handle

    cfs = np.zeros(x.ncols)
    kind, h = pc_handle(pc)
    call("famg_smooth_vector_pc_dev", _mat_of(mat)._h, kind, h, int(iterations), x._h, _f(cfs))
    if report:
        print("~||E||_A: " + " ".join(f"{c:.3f}" for c in cfs))
    return list(cfs)


def smooth_vector(mat, pc, iterations: int, near_null_dim: int, report: bool = False,
                  x0=None, seed: Optional[int] = None) -> Tuple[np.ndarray, List[float]]:
    """adaptivity.rs:307-390.  The reference draws the start block from an unseeded StandardNormal
    stream (``:321-329``); ``x0`` (n x near_null_dim) or ``seed`` pins it here."""
    m = _mat_of(mat)
    if x0 is None:
        x0 = np.random.default_rng(seed).standard_normal((m.nrows, near_null_dim))
    x = DeviceMat.from_host(m.ctx, x0)
    assert x.ncols == near_null_dim
    cfs = smooth_vector_dev(m, pc, iterations, x, report)
    return x.to_host(), cfs


def create_weights(nn_basis, mat) -> List[float]:
    """adaptivity.rs:434-443: 1 / (v^T A v) per column."""
    m = _mat_of(mat)
    v = DeviceMat.from_host(m.ctx, nn_basis)
    av = DeviceMat(m.ctx, v.nrows, v.ncols)
    m.apply_dev(av, v)
    vtav = np.zeros(v.ncols)
    call("famg_vec_coldot", v._h, av._h, _f(vtav))
    return list(1.0 / vtav)


def find_near_null(mat: SparseMatOp, iterations: int, near_null_dim: int, smoothing_block_size: float,
                   partitioner: Optional[Callable] = None, x0=None, seed: Optional[int] = None) -> np.ndarray:
    """adaptivity.rs:264-305: smooth with L1, build a block smoother over aggregates of about
    ``smoothing_block_size`` nodes from the first basis, smooth again with it.
    ``partitioner(op, near_null, weights, coarsening_factor) -> Partition`` stands where the
    reference calls ``PartitionerConfig{coarsening_factor, max_improvement_iters: 50}.build`` inside
    ``BlockSmootherConfig::build`` (``:279-289``); default: :class:`partitioners.PartitionerConfig`."""
    from .preconditioners.block_smoothers import BlockSmoother

    simple_pc = new_l1(mat.mat_ref())
    smooth_basis, _ = smooth_vector(mat, simple_pc, iterations, near_null_dim, False, x0=x0, seed=seed)
    weights = create_weights(smooth_basis, mat)
    if partitioner is None:
        from .partitioners import PartitionerConfig

        cfg = PartitionerConfig(coarsening_factor=smoothing_block_size, max_improvement_iters=50)
        partition = cfg.build_partition(mat, smooth_basis, weights)
    else:
        partition = partitioner(mat, smooth_basis, weights, smoothing_block_size)
    block_pc = BlockSmoother.new(mat, partition)
    x1 = None if x0 is None else np.asarray(x0)
    seed1 = None if seed is None else seed + 1
    smooth_basis, cfs = smooth_vector(mat, block_pc, iterations, near_null_dim, False, x0=x1, seed=seed1)
    print("||Ev||_A^(1/cycles): " + " ".join(f"{c:.2f}" for c in cfs))
    return smooth_basis


class AdaptiveConfig:
    """adaptivity.rs:28-165: build a :class:`Composite` of up to ``max_components`` multigrids, each
    from a hierarchy whose near-null block is what the composite built so far fails to damp.
    ``hierarchy_config`` / ``multigrid_config``: the mirrors in this package; the block smoother's
    coarsening factor (``multigrid_config.smoother_config.partitioner_config.coarsening_factor`` in the
    reference, ``:63-66``) is ``smoothing_block_size`` here."""

    def __init__(self, hierarchy_config, multigrid_config, target_convergence: Optional[float] = None, max_components: int = 5,
                 test_iters: int = 50, coarsening_near_null_dim: int = 32, include_constant_first_near_null: bool = True,
                 smoothing_block_size: float = 8.0, seed: Optional[int] = None):
        self.hierarchy_config, self.multigrid_config = hierarchy_config, multigrid_config
        self.target_convergence, self.max_components, self.test_iters = target_convergence, max_components, test_iters
        self.coarsening_near_null_dim = coarsening_near_null_dim
        self.include_constant_first_near_null = include_constant_first_near_null
        self.smoothing_block_size, self.seed = smoothing_block_size, seed

    def build(self, mat: SparseMatOp):
        from .hierarchy import thin_q
        from .preconditioners.composite import Composite

        k = self.coarsening_near_null_dim
        nn = find_near_null(mat, self.test_iters, k - 1, self.smoothing_block_size, seed=self.seed)       # :58-67
        nn_with_constant = np.ones((nn.shape[0], k), order="F")
        nn_with_constant[:, 1:] = nn                                                                        # :68-71
        basis = thin_q(nn_with_constant)
        weights = create_weights(basis, mat)                                                                # :73
        first = self.multigrid_config.build(self.hierarchy_config.build(mat, basis, weights))               # :110-115
        composite = Composite.new(mat.dyn_op(), first)
        for n_components in range(1, self.max_components):                                                  # :117-163
            seed = None if self.seed is None else self.seed + 100 * n_components
            smoothed, cfs = smooth_vector(mat, composite.clone(), self.test_iters // (2 * n_components - 1), k, False, seed=seed)
            n_vcycles = float(2 * n_components - 1)
            print("||Ev||_A^(1/cycles): " + " ".join(f"{c ** (1.0 / n_vcycles):.2f}" for c in cfs))
            composite.push(self.multigrid_config.build(self.hierarchy_config.build(mat, smoothed, cfs)))    # :155-161
        return composite
