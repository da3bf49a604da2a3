"""Aggregation interpolation and the Galerkin product (``src/interpolation/mod.rs``, aggregation
part: :28-157, :730-836, :927-946).  The classical / least-squares branch (:159-728) is host
combinatorics outside the hot path; it ends in the same ``R * (A * P)`` that
:func:`galerkin_product` provides."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from ._ffi import call, u64p, vp
from .core import SparseMatOp, SparseRowMat, _f, as_colmajor
from .partitioners import Partition, PartitionerConfig

JACOBI_WEIGHT = 0.66  # interpolation/mod.rs:814


@dataclass
class GalerkinCoarse:
    """interpolation/mod.rs:34-40."""
    interpolation: SparseRowMat
    restriction: SparseRowMat
    coarse_mat: SparseRowMat
    coarse_nn: np.ndarray
    partition: Partition


def tentative_prolongator(ctx, n_fine: int, partition: Partition, near_null, candidate_dimension: int = 1,
                          block_size: int = 1):
    """interpolation/mod.rs:747-809: per-aggregate thin SVD of the near-null block (host, tiny),
    P uploaded as device CSR.  Returns (P, coarse_near_null)."""
    nn = as_colmajor(near_null)
    # Partition keeps non-negative int64 arrays: same bits as the ABI's usize, passed without a copy
    ap = np.ascontiguousarray(partition.agg_ptr, dtype=np.int64)
    an = np.ascontiguousarray(partition.agg_nodes, dtype=np.int64)
    coarse_nn = np.zeros((partition.naggs() * candidate_dimension, nn.shape[1]), order="F")
    h = vp()
    call("famg_tentative_p", ctx._h, n_fine, block_size, nn.shape[1], candidate_dimension, _f(nn), max(nn.shape[0], 1),
         partition.naggs(), ap.ctypes.data_as(u64p), an.ctypes.data_as(u64p), C.byref(h), _f(coarse_nn))
    return SparseRowMat(ctx, h), coarse_nn


def tentative_prolongator_dev(partition, near_null_dev):
    """The same for a scalar problem with one near-null vector, on the device (``famg_tentative_p_dev``): device
    aggregates, device near-null column; returns (P, coarse near-null as a DeviceMat).  Bit-identical to the host path."""
    from .core import DeviceMat
    ctx = partition.ctx
    cnn = DeviceMat(ctx, partition.naggs(), 1)
    h = vp()
    call("famg_tentative_p_dev", partition._h, near_null_dev._h, C.byref(h), cnn._h)
    return SparseRowMat(ctx, h), cnn


def smooth_interpolation(mat: SparseRowMat, p: SparseRowMat, jacobi_weight: float = JACOBI_WEIGHT) -> SparseRowMat:
    """interpolation/mod.rs:927-946: (I - w D^-1 A) P, one SpGEMM with a fused epilogue."""
    h = vp()
    call("famg_smooth_interpolation", mat._h, p._h, float(jacobi_weight), C.byref(h))
    return SparseRowMat(mat.ctx, h)


def block_jacobi(mat: SparseRowMat, block_size: int, p: SparseRowMat) -> SparseRowMat:
    """interpolation/mod.rs:963-1028: P - 0.66 D_b^-1 A P with D_b^-1 from the per-block symmetric
    eigen-decomposition of A's block diagonal (two SpGEMMs, the ``+ P`` fused into the second)."""
    h = vp()
    call("famg_block_jacobi", mat._h, int(block_size), p._h, C.byref(h))
    return SparseRowMat(mat.ctx, h)


def smooth_p(mat: SparseRowMat, m_inv: SparseRowMat, p: SparseRowMat) -> SparseRowMat:
    """interpolation/mod.rs:1030-1040: ``m_inv * (-(mat * p)) + p``."""
    h = vp()
    call("famg_smooth_p", mat._h, m_inv._h, p._h, C.byref(h))
    return SparseRowMat(mat.ctx, h)


def galerkin_product(mat: SparseRowMat, p0: SparseRowMat, smoothing_steps: int = 1, jacobi_weight: float = JACOBI_WEIGHT,
                     block_size: int = 1):
    """interpolation/mod.rs:811-828: P = smooth^steps(P0) (``smooth_interpolation`` for block_size 1,
    ``block_jacobi`` otherwise); R = P^T; A_c = R (A P).  Returns (P, R, A_c)."""
    p, r, ac = vp(), vp(), vp()
    call("famg_galerkin_block", mat._h, p0._h, int(block_size), smoothing_steps, float(jacobi_weight), C.byref(p), C.byref(r),
         C.byref(ac))
    return SparseRowMat(mat.ctx, p), SparseRowMat(mat.ctx, r), SparseRowMat(mat.ctx, ac)


def smoothed_aggregation(fine_mat: SparseRowMat, partition: Partition, block_size: int, near_null,
                         candidate_dimension: int, smoothing_steps: int):
    """interpolation/mod.rs:730-836 -> (coarse_near_null, R, P, A_c, partition)."""
    n_fine = fine_mat.nrows
    assert n_fine % block_size == 0 and n_fine == partition.nnodes() * block_size  # :743-745
    p0, coarse_nn = tentative_prolongator(fine_mat.ctx, n_fine, partition, near_null, candidate_dimension, block_size)
    p, r, ac = galerkin_product(fine_mat, p0, smoothing_steps, block_size=block_size)
    return coarse_nn, r, p, ac, partition


class AggregationConfig:
    """interpolation/mod.rs:62-157.  ``partitioner``: a :class:`PartitionerConfig` (the reference's
    ``partitioner_config`` field; host-side algebraic aggregation) or a callable
    ``(level, op, near_null) -> Partition`` (e.g. :class:`GeometricPartitioner`)."""

    def __init__(self, smoothing_steps: int = 1, candidate_dimension: int = 4,
                 partitioner: Optional[Callable] = None):
        self.smoothing_steps = smoothing_steps
        self.candidate_dimension = candidate_dimension
        self.partitioner = partitioner

    @classmethod
    def new_unsmoothed(cls, partitioner, candidate_dimension: int) -> "AggregationConfig":
        return cls(0, candidate_dimension, partitioner)

    def device_path(self, op: SparseMatOp, near_null) -> bool:
        """Scalar problem, one near-null vector, a partitioner that can produce its aggregates on the device: the
        tentative prolongator is built there (no host pass over the fine grid, nothing uploaded)."""
        nn = as_colmajor(near_null) if not hasattr(near_null, "_h") else None
        k = near_null.ncols if nn is None else nn.shape[1]
        return (hasattr(self.partitioner, "device") and self.candidate_dimension == 1 and op.block_size() == 1 and k == 1)

    def build_dev(self, op: SparseMatOp, near_null_dev, level: int = 0):
        """``build`` on the device path: returns (GalerkinCoarse with coarse_nn = None, coarse near-null DeviceMat)."""
        partition = self.partitioner.device(level, op.mat_ref().ctx)
        assert op.mat_ref().nrows == partition.nnodes()  # interpolation/mod.rs:743-745
        p0, cnn_dev = tentative_prolongator_dev(partition, near_null_dev)
        p, r, ac = galerkin_product(op.mat_ref(), p0, self.smoothing_steps)
        return GalerkinCoarse(p, r, ac, None, partition), cnn_dev

    def build(self, op: SparseMatOp, near_null, nn_weights=None, level: int = 0) -> GalerkinCoarse:
        if self.partitioner is None:
            raise ValueError("AggregationConfig needs a partitioner (PartitionerConfig or callable)")
        if isinstance(self.partitioner, PartitionerConfig):
            # interpolation/mod.rs:135-138
            ratio = self.candidate_dimension / op.block_size()
            if nn_weights is None:
                raise ValueError("the algebraic partitioner needs nn_weights (one per near-null vector)")
            partition = self.partitioner.scaled(ratio).build_partition(op, near_null, nn_weights)
        else:
            partition = self.partitioner(level, op, near_null)
        coarse_nn, r, p, ac, partition = smoothed_aggregation(op.mat_ref(), partition, op.block_size(), near_null,
                                                              self.candidate_dimension, self.smoothing_steps)
        return GalerkinCoarse(p, r, ac, coarse_nn, partition)


InterpolationConfig = AggregationConfig  # the Aggregation variant of interpolation/mod.rs:28-32
