"""Small callers of the hot path from ``src/utils.rs`` and ``multigrid.rs``: ``mats_are_equal``
(utils.rs:32-58), ``approx_convergence_factor`` (utils.rs:691-736) and ``symmetry_test``
(multigrid.rs:520-580).  They are written over the ``LinOp::apply`` surface every operator of this
package has (``apply(ndarray) -> ndarray``, host buffers staged through the device), so the logic is
also testable with plain numpy operators."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def mats_are_equal(left, right, tol: float = 1e-12) -> bool:
    """utils.rs:32-58: same shape and nnz, identical (row, col) sequence, every value within ``tol``
    absolutely and relatively.  (The reference divides by ``left.val.max(right.val)``, which is
    negative for a pair of negative values -- any difference then passes; max(|l|, |r|) is used here.)"""
    if left.shape != right.shape or left.compute_nnz() != right.compute_nnz():
        return False
    lrp, lci, lv = left.to_host()
    rrp, rci, rv = right.to_host()
    if not (np.array_equal(lrp, rrp) and np.array_equal(lci, rci)):
        return False
    absd = np.abs(lv - rv)
    den = np.maximum(np.abs(lv), np.abs(rv))
    rel = np.divide(absd, den, out=np.zeros_like(absd), where=den > 0)
    return bool(np.all(absd <= tol) and np.all(rel <= tol))


def approx_convergence_factor(mat, pc, iterations: int = 100, test_vecs: int = 5, seed: Optional[int] = None) -> float:
    """utils.rs:691-736: power iteration on ``E = I - M^-1 A`` with A-norm normalisation of every column
    before each step; returns the mean A-norm of the columns after the last step (an estimate of
    ``||E||_A``).  ``mat`` / ``pc``: anything with ``apply(ndarray) -> ndarray`` (``SparseMatOp`` is
    unwrapped through ``dyn_op()``).  The reference draws the start block from an unseeded stream."""
    op = mat.dyn_op() if hasattr(mat, "dyn_op") else mat
    n = op.nrows() if callable(getattr(op, "nrows", None)) else op.nrows
    x = np.asfortranarray(np.random.default_rng(seed).standard_normal((n, test_vecs)))

    def a_norms(w):
        return np.sqrt(np.einsum("ij,ij->j", w, op.apply(w)))

    for _ in range(iterations):
        x = x / a_norms(x)                      # normalize_cols (:719-724)
        x = x - pc.apply(op.apply(x))           # ErrorPropogator::apply (adaptivity.rs:191-198)
    return float(np.sum(a_norms(x)) / test_vecs)


def symmetry_test(bilinear_op, test_dim: int = 20, seed: Optional[int] = None) -> Tuple[float, float]:
    """multigrid.rs:520-580: max absolute / relative difference between ``u^T (B v)`` and ``(B^T u)^T v``
    over ``test_dim`` random pairs; prints and returns both."""
    n = bilinear_op.nrows() if callable(getattr(bilinear_op, "nrows", None)) else bilinear_op.nrows
    rng = np.random.default_rng(seed)
    u, v = rng.standard_normal((n, test_dim)), rng.standard_normal((n, test_dim))
    utav = u.T @ bilinear_op.apply(v)
    t_apply = getattr(bilinear_op, "transpose_apply", bilinear_op.apply)
    maybe_utav = (v.T @ t_apply(u)).T
    err = np.abs(utav - maybe_utav)
    den = np.maximum(utav, maybe_utav)          # as the reference (:566), sign-sensitive
    rel = np.abs(np.divide(err, den, out=np.zeros_like(err), where=den != 0))
    print(f"max error (abs , rel): {err.max():.2e} , {rel.max():.2e}")
    return float(err.max()), float(rel.max())
