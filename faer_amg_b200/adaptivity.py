"""The one piece of ``src/adaptivity.rs`` that sits on the hot path: ``ErrorPropogator`` (sic),
the operator ``E = I - M^-1 A`` whose ``apply`` (adaptivity.rs:191-198) is the inner loop of the
near-null search (``:351-354``) and of compatible relaxation (``interpolation/mod.rs:623-625``).
The search drivers themselves (``find_near_null``, ``smooth_vector``, ``AdaptiveConfig``) are host
logic outside the scope of this package."""
from __future__ import annotations

import numpy as np

from .core import DeviceMat, ParSpmmOp, SparseMatOp, SparseRowMat
from .preconditioners.multigrid import Multigrid


class ErrorPropogator:
    """``out = x - M^-1 (A x)`` on n x k blocks, device-resident (SpMM of width k + precond apply)."""

    def __init__(self, op, pc):
        self.op = op.mat_ref() if isinstance(op, SparseMatOp) else op.mat if isinstance(op, ParSpmmOp) else op
        self.pc = pc

    def nrows(self) -> int:
        return self.op.nrows

    ncols = nrows

    def apply_dev(self, out: DeviceMat, x: DeviceMat, work: DeviceMat):
        self.op.apply_dev(work, x)          # work = A x
        self.pc.apply_dev(out, work)        # out = M^-1 work
        out.axpby(1.0, x, -1.0)             # out = x - out

    def apply(self, x) -> np.ndarray:
        ctx = self.op.ctx
        X = DeviceMat.from_host(ctx, x)
        out, work = DeviceMat(ctx, X.nrows, X.ncols), DeviceMat(ctx, X.nrows, X.ncols)
        self.apply_dev(out, X, work)
        return out.to_host()
