"""Level structure (``src/hierarchy.rs``)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np

from ._ffi import call
from .core import DeviceMat, SparseMatOp, SparseRowMat, _f, as_colmajor
from .interpolation import AggregationConfig
from .partitioners import Partition
from .preconditioners.smoothers import StationaryIteration, new_l1


def thin_q(m) -> np.ndarray:
    """``qr().compute_thin_Q()`` (hierarchy.rs:228), R with a positive diagonal."""
    m = as_colmajor(m).copy(order="F")
    call("famg_thin_q", m.shape[0], m.shape[1], _f(m), max(m.shape[0], 1))
    return m


def thin_q_dev(x: DeviceMat):
    """Device-resident thin Q (CholeskyQR2, k <= 64), in place."""
    call("famg_thin_q_dev", x._h)


class HierarchyConfig:
    """hierarchy.rs:22-53."""

    def __init__(self, coarsest_dim: int = 1000, interpolation_config: Optional[AggregationConfig] = None,
                 max_levels: Optional[int] = None):
        self.coarsest_dim = coarsest_dim
        self.interpolation_config = interpolation_config
        self.max_levels = max_levels

    def build(self, base_matrix: SparseMatOp, near_null, nn_weights=None) -> "Hierarchy":
        h = Hierarchy(base_matrix, near_null, nn_weights, self)
        h.coarsen()
        return h


class Hierarchy:
    """hierarchy.rs:61-360: operators / restrictions / interpolations / partitions / near_nulls."""

    def __init__(self, fine_op: SparseMatOp, fine_near_null, nn_weights, config: HierarchyConfig):
        self._operators: List[SparseMatOp] = [fine_op]
        self._restrictions: List[SparseRowMat] = []
        self._interpolations: List[SparseRowMat] = []
        self._partitions: List[Partition] = []
        self._near_nulls: List[np.ndarray] = [as_colmajor(fine_near_null)]
        self._nn_weights = [nn_weights]
        self.config = config

    def coarsen(self):
        """hierarchy.rs:190-248."""
        cfg = self.config
        max_levels = cfg.max_levels if cfg.max_levels is not None else 1 << 62
        level, coarse_dim = 1, 1 << 62
        nn_dev_prev = None  # the current level's near-null column on the device (device path: levels >= 1 never leave it)
        while coarse_dim > cfg.coarsest_dim and level < max_levels:
            fine = self.current_op()
            near_null = self._near_nulls[-1]
            icfg = cfg.interpolation_config
            if hasattr(icfg, "device_path") and icfg.device_path(fine, near_null):
                if nn_dev_prev is None:
                    nn_dev_prev = DeviceMat.from_host(fine.mat_ref().ctx, near_null)
                g, nn_dev = icfg.build_dev(fine, nn_dev_prev, level=level - 1)
            else:
                g = icfg.build(fine, near_null, self._nn_weights[-1], level=level - 1)
                nn_dev = DeviceMat.from_host(g.coarse_mat.ctx, g.coarse_nn)
            coarse_op = SparseMatOp(g.coarse_mat, cfg.interpolation_config.candidate_dimension)
            coarse_dim = coarse_op.mat_ref().nrows
            # smooth the coarse near-null: 3-step L1 stationary iteration (:217-226), thin QR (:228)
            l1 = new_l1(coarse_op.mat_ref())
            StationaryIteration(coarse_op.mat_ref(), l1, 3).apply_in_place_dev(nn_dev)
            coarse_nn = thin_q(nn_dev.to_host())
            nn_dev.upload(coarse_nn)
            nn_dev_prev = nn_dev
            self.add_level(coarse_op, g.partition, coarse_nn, g.interpolation, g.restriction)
            level += 1

    def add_level(self, coarse_op: SparseMatOp, partition, near_null, interpolation: SparseRowMat, restriction: SparseRowMat):
        """hierarchy.rs:250-271 (same shape asserts)."""
        assert interpolation.nrows == restriction.ncols
        assert interpolation.nrows == self.current_mat_ref().nrows
        assert interpolation.ncols == restriction.nrows
        assert interpolation.ncols == coarse_op.mat_ref().ncols
        self._operators.append(coarse_op)
        self._partitions.append(partition)
        self._restrictions.append(restriction)
        self._interpolations.append(interpolation)
        self._near_nulls.append(as_colmajor(near_null))

    def get_config(self):
        return self.config

    def levels(self) -> int:
        return len(self._operators)

    def operators(self):
        return self._operators

    def get_op(self, level: int) -> SparseMatOp:
        return self._operators[level]

    def get_arc_mat(self, level: int) -> SparseRowMat:
        return self._operators[level].arc_mat()

    get_mat_ref = get_arc_mat

    def current_op(self) -> SparseMatOp:
        return self._operators[-1]

    def current_mat_ref(self) -> SparseRowMat:
        return self._operators[-1].mat_ref()

    current_arc_mat = current_mat_ref

    def partitions(self):
        return self._partitions

    def get_partition(self, level: int) -> Partition:
        return self._partitions[level]

    def restrictions(self):
        return self._restrictions

    def get_restriction(self, level: int) -> SparseRowMat:
        return self._restrictions[level]

    def interpolations(self):
        return self._interpolations

    def get_interpolation(self, level: int) -> SparseRowMat:
        return self._interpolations[level]

    def near_nulls(self):
        return self._near_nulls

    def get_near_null(self, level: int) -> np.ndarray:
        return self._near_nulls[level]

    def get_nn_weights(self, level: int):
        return self._nn_weights[min(level, len(self._nn_weights) - 1)]

    def grid_complexity(self) -> float:  # hierarchy.rs:346-350
        return sum(op.mat_ref().nrows for op in self._operators) / self._operators[0].mat_ref().nrows

    def op_complexity(self) -> float:  # hierarchy.rs:352-360
        return sum(op.mat_ref().compute_nnz() for op in self._operators) / self._operators[0].mat_ref().compute_nnz()

    def __repr__(self):  # hierarchy.rs:121-170 (stats table)
        rows = [f"{'lev':>4} {'ndofs':>12} {'nnz':>14} {'nnz/row':>8}"]
        for i, op in enumerate(self._operators):
            m = op.mat_ref()
            rows.append(f"{i:>4} {m.nrows:>12} {m.nnz:>14} {m.nnz / max(m.nrows, 1):>8.2f}")
        rows.append(f"operator complexity {self.op_complexity():.3f}, grid complexity {self.grid_complexity():.3f}")
        return "\n".join(rows)
