"""Coarsest-level exact solve (``src/preconditioners/coarse_solvers.rs``)."""
from __future__ import annotations

import ctypes as C

from .._ffi import call, vp
from ..core import SparseRowMat
from .smoothers import Smoother


class SparseCholeskySolve(Smoother):
    """coarse_solvers.rs:166-276: x = A^-1 f. Host Cholesky at build, device GEMV with the explicit
    inverse at apply (csrc/coarse.cu)."""

    @classmethod
    def new(cls, sym_mat: SparseRowMat) -> "SparseCholeskySolve":
        h = vp()
        call("famg_smoother_cholesky", sym_mat._h, C.byref(h))
        return cls(sym_mat.ctx, h)


class CoarseSolverKind:
    """coarse_solvers.rs:14-42: only Cholesky is implemented upstream."""
    Cholesky = "cholesky"
    Svd = "svd"
    Eigh = "eigh"

    @staticmethod
    def build_from_sparse(kind: str, mat: SparseRowMat) -> Smoother:
        if kind != CoarseSolverKind.Cholesky:
            raise NotImplementedError(kind)  # coarse_solvers.rs:26,30
        return SparseCholeskySolve.new(mat)
