"""Block smoother over aggregates (``src/preconditioners/block_smoothers.rs``) -- tier 2."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .._ffi import call, u64p, vp
from ..core import SparseMatOp
from ..partitioners import Partition
from .smoothers import Smoother


class BlockSmoother(Smoother):
    """block_smoothers.rs:78-214: one diagonally-compensated block (``:293-324``) per aggregate,
    solved exactly; materialised as the block-diagonal M^-1 of ``into_sparse_mat`` (``:125-146``)
    and applied as one SpMV."""

    @classmethod
    def new(cls, op: SparseMatOp, partition: Partition) -> "BlockSmoother":
        ap = np.ascontiguousarray(partition.agg_ptr, dtype=np.int64)   # same bits as usize, no copy
        an = np.ascontiguousarray(partition.agg_nodes, dtype=np.int64)
        h = vp()
        # vdim == 1: diagonally_compensate (:293-324); vdim > 1: diagonally_compensate_vector (:326-400)
        call("famg_smoother_block_vector", op.mat_ref()._h, op.block_size(), partition.naggs(), ap.ctypes.data_as(u64p),
             an.ctypes.data_as(u64p), C.byref(h))
        s = cls(op.mat_ref().ctx, h)
        s.partition = partition
        return s


class BlockSmootherConfig:
    """block_smoothers.rs:36-76 (BlockSolver(Cholesky) only, like upstream)."""

    def __init__(self, partitioner=None):
        self.partitioner = partitioner

    def build_from_partition(self, op: SparseMatOp, partition: Partition) -> BlockSmoother:
        return BlockSmoother.new(op, partition)
