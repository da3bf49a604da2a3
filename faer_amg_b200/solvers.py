"""PCG and the stationary solver as the reference drives them (``src/utils.rs:553-661``,
``examples/simple_geometric.rs:117-174, 229-267``)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from . import _ffi
from ._ffi import CgInfoStruct, FamgError, call
from .core import DeviceMat, ParSpmmOp, SparseMatOp, SparseRowMat, _f
from .preconditioners.multigrid import Multigrid
from .preconditioners.smoothers import Smoother

PC_NONE, PC_SMOOTHER, PC_MG = 0, 1, 2


@dataclass
class CgParams:
    """faer CgParams as filled by utils.rs:580-586 / simple_geometric.rs:229-234."""
    abs_tolerance: float = 0.0
    rel_tolerance: float = 1e-12
    max_iters: int = 1000
    initial_guess_zero: bool = True


@dataclass
class CgInfo:
    abs_residual: float
    rel_residual: float
    iter_count: int


class CgError(RuntimeError):
    """CgError::NoConvergence / non-positive-definite variants (utils.rs:647-659)."""

    def __init__(self, kind: str, abs_residual: float, rel_residual: float, iter_count: int):
        super().__init__(f"{kind}: abs {abs_residual:.3e}, rel {rel_residual:.3e} after {iter_count} iterations")
        self.kind, self.abs_residual, self.rel_residual, self.iter_count = kind, abs_residual, rel_residual, iter_count


def _mat(op) -> SparseRowMat:
    if isinstance(op, SparseMatOp):
        return op.mat_ref()
    return op.mat if isinstance(op, ParSpmmOp) else op


def _pc(precond):
    from .preconditioners.composite import pc_handle

    return pc_handle(precond)


def _finish(fn_status, info, max_iters):
    out = CgInfo(info.abs_residual, info.rel_residual, int(info.iter_count))
    if fn_status == _ffi.OK:
        return out
    if fn_status == _ffi.ERR_NO_CONVERGENCE:
        raise CgError("NoConvergence", out.abs_residual, out.rel_residual, out.iter_count)
    if fn_status == _ffi.ERR_NOT_SPD:
        raise CgError("NonPositiveDefinite", out.abs_residual, out.rel_residual, out.iter_count)
    _ffi.check(fn_status)


def conjugate_gradient(dst: np.ndarray, precond, mat, rhs, params: CgParams) -> CgInfo:
    """faer ``conjugate_gradient(out, precond, mat, rhs, params, ..)`` with host buffers:
    ``dst`` holds the initial guess on entry (unless ``initial_guess_zero``) and the solution on
    return.  Raises :class:`CgError` like the reference returns ``Err``."""
    m = _mat(mat)
    b = np.ascontiguousarray(np.asarray(rhs, dtype=np.float64).reshape(-1))
    assert dst.dtype == np.float64 and (dst.flags.c_contiguous or dst.flags.f_contiguous)
    x = dst.reshape(-1)
    kind, h = _pc(precond)
    info = CgInfoStruct()
    st = _ffi.lib().famg_pcg_solve(m._h, kind, h, _f(x), _f(b), params.rel_tolerance, params.abs_tolerance,
                                    params.max_iters, 1 if params.initial_guess_zero else 0, C.byref(info))
    return _finish(st, info, params.max_iters)


def conjugate_gradient_dev(dst: DeviceMat, precond, mat, rhs: DeviceMat, params: CgParams) -> CgInfo:
    """Device-resident variant: nothing crosses PCIe except three scalars per iteration."""
    m = _mat(mat)
    kind, h = _pc(precond)
    info = CgInfoStruct()
    st = _ffi.lib().famg_pcg_solve_dev(m._h, kind, h, dst._h, rhs._h, params.rel_tolerance, params.abs_tolerance,
                                        params.max_iters, 1 if params.initial_guess_zero else 0, C.byref(info))
    return _finish(st, info, params.max_iters)


def stationary_solver(x: np.ndarray, b, op, pc, max_iter: int, rel_tolerance: float) -> int:
    """examples/simple_geometric.rs:117-158; returns the iteration count, x updated in place."""
    m = _mat(op)
    bb = np.ascontiguousarray(np.asarray(b, dtype=np.float64).reshape(-1))
    kind, h = _pc(pc)
    iters = C.c_int64()
    call("famg_stationary_solve", m._h, kind, h, _f(x.reshape(-1)), _f(bb), rel_tolerance, max_iter, C.byref(iters))
    return iters.value


def test_solver(op, pc, initial_guess: Optional[np.ndarray], rhs: Optional[np.ndarray], max_iters: int,
                tolerance: float) -> Tuple[int, float, np.ndarray]:
    """The PCG half of utils.rs:553-633 (abs_tolerance 0, rel_tolerance = tolerance); like
    report_cg (:635-661) a failed solve reports (1000 iterations, last rel residual)."""
    m = _mat(op)
    b = np.zeros(m.nrows) if rhs is None else np.asarray(rhs, dtype=np.float64).reshape(-1)
    x = np.zeros(m.nrows) if initial_guess is None else np.array(initial_guess, dtype=np.float64).reshape(-1)
    params = CgParams(0.0, tolerance, max_iters, initial_guess is None)
    try:
        info = conjugate_gradient(x, pc, m, b, params)
        return info.iter_count, info.rel_residual, x
    except CgError as e:
        return 1000, e.rel_residual, x


test_solver.__test__ = False  # not a pytest test
