"""``Multigrid`` and ``MultigridConfig`` (``src/preconditioners/multigrid.rs``)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from .. import _ffi
from .._ffi import call, vp
from ..core import DeviceMat, ParSpmmOp, SparseRowMat, _f, as_colmajor
from .coarse_solvers import CoarseSolverKind
from .smoothers import Smoother, new_jacobi, new_l1, new_l2


def _mat(op) -> SparseRowMat:
    return op.mat if isinstance(op, ParSpmmOp) else op


class Multigrid:
    """multigrid.rs:171-249 + the cycle (:251-380) as a device-resident, CUDA-graph-replayed
    preconditioner.  ``apply`` is ``LinOp::apply`` (one mu-cycle from a zero guess); transpose ==
    apply (symmetric multigrid only, :492-503)."""

    def __init__(self, op, smoother: Smoother):
        m = _mat(op)
        h = vp()
        call("famg_mg_create", m._h, smoother._h, C.byref(h))
        self._h, self.ctx = h, m.ctx
        self._keep = [m, smoother]
        self._n = m.nrows
        self._cycle_type, self._smoothing_steps = 1, 1

    new = classmethod(lambda cls, op, smoother: cls(op, smoother))

    def with_cycle_type(self, mu: int) -> "Multigrid":
        assert mu > 0  # multigrid.rs:204
        self._cycle_type = mu
        call("famg_mg_set_cycle", self._h, self._cycle_type, self._smoothing_steps)
        return self

    def with_smoothing_steps(self, steps: int) -> "Multigrid":
        assert steps > 0  # multigrid.rs:210
        self._smoothing_steps = steps
        call("famg_mg_set_cycle", self._h, self._cycle_type, self._smoothing_steps)
        return self

    def add_level(self, op, smoother: Smoother, r, p):
        """multigrid.rs:228-239; shapes checked as in hierarchy.rs:258-264."""
        m, rr, pp = _mat(op), _mat(r), _mat(p)
        call("famg_mg_add_level", self._h, m._h, smoother._h, rr._h, pp._h)
        self._keep += [m, smoother, rr, pp]

    def levels(self) -> int:
        n = C.c_int()
        call("famg_mg_levels", self._h, C.byref(n))
        return n.value

    def cycle_type(self) -> int:
        return self._cycle_type

    def nrows(self) -> int:
        return self._n

    ncols = nrows

    def apply(self, rhs) -> np.ndarray:
        rhs = as_colmajor(rhs)
        out = np.empty_like(rhs, order="F")
        call("famg_mg_apply", self._h, _f(out), max(self._n, 1), _f(rhs), max(self._n, 1), rhs.shape[1])
        return out

    transpose_apply = apply
    conj_apply = apply

    def apply_dev(self, out: DeviceMat, rhs: DeviceMat):
        call("famg_mg_apply_dev", self._h, out._h, rhs._h)

    def cycle_bytes(self, k: int = 1) -> float:
        b = C.c_double()
        call("famg_mg_cycle_bytes", self._h, k, C.byref(b))
        return b.value

    def __del__(self):
        try:
            _ffi.lib().famg_mg_destroy(self._h)
        except Exception:
            pass


class MultigridConfig:
    """multigrid.rs:27-164.  ``smoother`` selects what every non-coarsest level gets:
    'block' (the reference default, BlockSmootherConfig -> needs the hierarchy's partitions),
    'l1' | 'l2' | 'jacobi' (the diagonal smoothers the reference wires by hand through
    ``Multigrid::add_level``, simple_geometric.rs:204-224)."""

    def __init__(self, mu: int = 1, smoothing_steps: int = 1, coarse_solver: Optional[str] = CoarseSolverKind.Cholesky,
                 smoother: str = "l1", omega: float = 0.66):
        self.mu, self.smoothing_steps, self.coarse_solver = mu, smoothing_steps, coarse_solver
        self.smoother, self.omega = smoother, omega

    def _level_smoother(self, hierarchy, level) -> Smoother:
        op = hierarchy.get_op(level)
        mat = op.mat_ref()
        if self.smoother == "l1":
            return new_l1(mat)
        if self.smoother == "l2":
            return new_l2(mat)
        if self.smoother == "jacobi":
            return new_jacobi(mat, self.omega)
        if self.smoother == "block":
            from .block_smoothers import BlockSmoother
            return BlockSmoother.new(op, hierarchy.get_partition(level))
        raise ValueError(self.smoother)

    def build(self, hierarchy, is_tail: bool = False) -> Multigrid:
        """``is_tail``: the hierarchy is the replicated tail of a distributed one -- its last level is the
        coarsest level of the whole hierarchy even when it is the tail's only level."""
        level_count = hierarchy.levels()
        smoothers = []
        for level in range(level_count):  # multigrid.rs:105-119
            if level + 1 == level_count and self.coarse_solver is not None and (level_count > 1 or is_tail):
                smoothers.append(CoarseSolverKind.build_from_sparse(self.coarse_solver, hierarchy.get_mat_ref(level)))
            else:
                smoothers.append(self._level_smoother(hierarchy, level))
        mg = Multigrid(hierarchy.get_op(0).dyn_op(), smoothers[0])
        for level in range(1, level_count):  # multigrid.rs:142-160 (F10b: the level's own operator)
            mg.add_level(hierarchy.get_op(level).dyn_op(), smoothers[level], hierarchy.get_restriction(level - 1),
                         hierarchy.get_interpolation(level - 1))
        return mg.with_cycle_type(self.mu).with_smoothing_steps(self.smoothing_steps)
