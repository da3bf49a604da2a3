"""Diagonal smoothers and the stationary iteration (``src/preconditioners/smoothers.rs``)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _ffi
from .._ffi import call, vp
from ..core import DeviceMat, SparseRowMat, _f, as_colmajor

DIAG_L1, DIAG_L2, DIAG_JACOBI = 0, 1, 2


class Smoother:
    """Base of everything a ``Multigrid`` level accepts as ``Arc<dyn BiPrecond<f64>>``."""

    def __init__(self, ctx, handle):
        self._h, self.ctx = handle, ctx
        n = C.c_int64()
        call("famg_smoother_dim", handle, C.byref(n))
        self.n = n.value

    def nrows(self):
        return self.n

    ncols = nrows

    def apply(self, rhs) -> np.ndarray:
        """``LinOp::apply``: out = M^-1 rhs (host buffers)."""
        rhs = as_colmajor(rhs)
        out = np.empty_like(rhs, order="F")
        call("famg_smoother_apply", self._h, _f(out), max(self.n, 1), _f(rhs), max(self.n, 1), rhs.shape[1])
        return out

    # symmetric smoothers only, like the reference's Diag / Cholesky / BlockSmoother
    transpose_apply = apply
    apply_in_place = apply

    def apply_dev(self, out: DeviceMat, rhs: DeviceMat):
        call("famg_smoother_apply_dev", self._h, out._h, rhs._h)

    def __del__(self):
        try:
            _ffi.lib().famg_smoother_destroy(self._h)
        except Exception:
            pass


class Diag(Smoother):
    """faer ``Diag<f64>`` used directly as the preconditioner (smoothers.rs:40-41)."""

    @classmethod
    def from_host(cls, ctx, d) -> "Diag":
        d = np.ascontiguousarray(d, dtype=np.float64)
        h = vp()
        call("famg_smoother_diag_from_host", ctx._h, len(d), _f(d), C.byref(h))
        return cls(ctx, h)

    def column_vector(self) -> np.ndarray:
        d = np.empty(self.n)
        call("famg_smoother_diag_download", self._h, _f(d))
        return d


def _diag(mat: SparseRowMat, kind: int, omega: float = 0.0) -> Diag:
    h = vp()
    call("famg_smoother_diag", mat._h, kind, float(omega), C.byref(h))
    return Diag(mat.ctx, h)


def new_l1(mat: SparseRowMat) -> Diag:
    """smoothers.rs:63-76."""
    return _diag(mat, DIAG_L1)


def new_l2(mat: SparseRowMat) -> Diag:
    """smoothers.rs:43-61."""
    return _diag(mat, DIAG_L2)


def new_jacobi(mat: SparseRowMat, omega: float) -> Diag:
    """smoothers.rs:78-86."""
    return _diag(mat, DIAG_JACOBI, omega)


class SmootherKind:
    """smoothers.rs:14-33. GaussSeidel / SymGaussSeidel are ``unimplemented!()`` upstream too."""

    def __init__(self, kind: str, omega: float = 0.66):
        self.kind, self.omega = kind, omega

    L1 = L2 = None  # filled below

    @staticmethod
    def Jacobi(omega: float) -> "SmootherKind":
        return SmootherKind("jacobi", omega)

    def build(self, mat: SparseRowMat) -> Diag:
        if self.kind == "l1":
            return new_l1(mat)
        if self.kind == "l2":
            return new_l2(mat)
        if self.kind == "jacobi":
            return new_jacobi(mat, self.omega)
        raise NotImplementedError(self.kind)  # smoothers.rs:26-27


SmootherKind.L1 = SmootherKind("l1")
SmootherKind.L2 = SmootherKind("l2")


def smooth(x: DeviceMat, b: DeviceMat, op: SparseRowMat, pc: Smoother, max_iter: int):
    """``smooth()`` (multigrid.rs:407-424): max_iter x { x += M^-1 (b - A x) } on the device."""
    call("famg_smooth_dev", op._h, pc._h, x._h, b._h, max_iter)


class StationaryIteration:
    """smoothers.rs:88-159 with a ``Diag`` preconditioner (its only use, hierarchy.rs:219-226)."""

    def __init__(self, mat: SparseRowMat, prec: Diag, iters: int):
        self.mat, self.prec, self.iters = mat, prec, iters

    def apply_in_place_dev(self, io: DeviceMat):
        call("famg_stationary_iteration_dev", self.mat._h, self.prec._h, self.iters, io._h)

    def apply(self, rhs) -> np.ndarray:
        io = DeviceMat.from_host(self.mat.ctx, rhs)
        self.apply_in_place_dev(io)
        return io.to_host()

    def transpose_apply(self, rhs) -> np.ndarray:
        """smoothers.rs:179-197: work1 = rhs; iters x { work2 = M work1; out = A work2; work1 -= out }."""
        ctx = self.mat.ctx
        w1 = DeviceMat.from_host(ctx, rhs)
        w2, out = DeviceMat(ctx, w1.nrows, w1.ncols), DeviceMat(ctx, w1.nrows, w1.ncols)
        for _ in range(self.iters):
            self.prec.apply_dev(w2, w1)
            self.mat.apply_dev(out, w2)
            w1.axpby(-1.0, out, 1.0)
        return w1.to_host()
