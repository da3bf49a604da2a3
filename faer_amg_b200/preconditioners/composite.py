"""``Composite`` (``src/preconditioners/composite.rs``): symmetric multiplicative combination of
several preconditioners around one operator, device-resident (``famg_composite_*``)."""
from __future__ import annotations

import ctypes as C
from typing import List

import numpy as np

from .. import _ffi
from .._ffi import call, vp
from ..core import DeviceMat, ParSpmmOp, SparseMatOp, SparseRowMat, as_colmajor

PC_NONE, PC_SMOOTHER, PC_MG, PC_COMPOSITE = 0, 1, 2, 3


def pc_handle(precond):
    """(pc_kind, handle) of any preconditioner object of this package -- the pair the C ABI takes."""
    from .multigrid import Multigrid
    from .smoothers import Smoother

    if precond is None:
        return PC_NONE, None
    if isinstance(precond, Multigrid):
        return PC_MG, precond._h
    if isinstance(precond, Smoother):
        return PC_SMOOTHER, precond._h
    if isinstance(precond, Composite):
        return PC_COMPOSITE, precond._h
    raise TypeError(f"unsupported preconditioner {type(precond)}")


class Composite:
    """composite.rs:11-100.  ``apply`` (:66-83): out = 0; ws = rhs; for every component in reverse
    order, then for components[1:] in order: ws <- c^-1 ws; out += ws; ws = rhs - A out."""

    def __init__(self, mat, first_component=None):
        m = mat.mat_ref() if isinstance(mat, SparseMatOp) else mat.mat if isinstance(mat, ParSpmmOp) else mat
        h = vp()
        call("famg_composite_create", m._h, C.byref(h))
        self._h, self.ctx, self._mat = h, m.ctx, m
        self._components: List = []
        if first_component is not None:
            self.push(first_component)

    @classmethod
    def new(cls, mat, first_component) -> "Composite":
        return cls(mat, first_component)

    @classmethod
    def new_with_components(cls, mat, components) -> "Composite":
        c = cls(mat)
        for comp in components:
            c.push(comp)
        return c

    def push(self, component):
        kind, h = pc_handle(component)
        call("famg_composite_push", self._h, kind, h)
        self._components.append(component)  # keeps the borrowed handle alive

    def components(self) -> List:
        return self._components

    def get_mat(self) -> SparseRowMat:
        return self._mat

    def nrows(self) -> int:
        return self._mat.nrows

    ncols = nrows

    def apply_dev(self, out: DeviceMat, rhs: DeviceMat):
        call("famg_composite_apply_dev", self._h, out._h, rhs._h)

    def apply(self, rhs) -> np.ndarray:
        x = DeviceMat.from_host(self.ctx, as_colmajor(rhs))
        out = DeviceMat(self.ctx, x.nrows, x.ncols)
        self.apply_dev(out, x)
        return out.to_host()

    conj_apply = apply

    def clone(self) -> "Composite":
        """``composite.clone()`` (adaptivity.rs:104): a snapshot of the current component list."""
        return Composite.new_with_components(self._mat, list(self._components))

    def __del__(self):
        try:
            _ffi.lib().famg_composite_destroy(self._h)
        except Exception:
            pass
