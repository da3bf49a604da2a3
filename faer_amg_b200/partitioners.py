"""``Partition`` container, deterministic geometric aggregation and the algebraic partitioner.

The reference's algebraic partitioner (``src/partitioners/``) is host-side graph work that the
north-star keeps out of the GPU path; it is also non-deterministic.  The hot path only
consumes its *output*, a ``Partition`` (``partitioners/mod.rs:23-27``: node_to_agg + agg_to_node).
This module provides that container, the structured-grid aggregates used by the benchmark
configurations, and ``PartitionerConfig`` -- the host restatement of
``PartitionerConfig::build_partition`` (least-squares strength graph, greedy modularity matching,
node-swap refinement; ``csrc/partition.cu``, host-only C++) with documented tie-breaking.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np


class Partition:
    """Aggregates in CSR form: ``agg_nodes[agg_ptr[a]:agg_ptr[a+1]]`` ascending (BTreeSet order)."""

    def __init__(self, agg_ptr, agg_nodes, nnodes: int, validate: bool = True):
        self.agg_ptr = np.ascontiguousarray(agg_ptr, dtype=np.int64)
        self.agg_nodes = np.ascontiguousarray(agg_nodes, dtype=np.int64)
        self._nnodes = int(nnodes)
        if validate:
            self.validate()

    @classmethod
    def from_node_to_agg(cls, node_to_agg) -> "Partition":
        node_to_agg = np.asarray(node_to_agg, dtype=np.int64)
        naggs = int(node_to_agg.max()) + 1 if len(node_to_agg) else 0
        order = np.argsort(node_to_agg, kind="stable")
        ptr = np.zeros(naggs + 1, dtype=np.int64)
        np.cumsum(np.bincount(node_to_agg, minlength=naggs), out=ptr[1:])
        return cls(ptr, order, len(node_to_agg))

    def validate(self):  # partitioners/mod.rs:144-154
        if self.agg_ptr[0] != 0 or self.agg_ptr[-1] != self._nnodes or len(self.agg_nodes) != self._nnodes:
            raise ValueError("partition does not cover every node exactly once")
        if self._nnodes and not np.array_equal(np.sort(self.agg_nodes), np.arange(self._nnodes)):
            raise ValueError("partition does not cover every node exactly once")
        if np.any(np.diff(self.agg_ptr) <= 0):
            raise ValueError("empty aggregate")

    def naggs(self) -> int:
        return len(self.agg_ptr) - 1

    def nnodes(self) -> int:
        return self._nnodes

    def aggregates(self):
        return [self.agg_nodes[self.agg_ptr[a]:self.agg_ptr[a + 1]] for a in range(self.naggs())]

    def node_to_agg(self) -> np.ndarray:
        out = np.empty(self._nnodes, dtype=np.int64)
        out[self.agg_nodes] = np.repeat(np.arange(self.naggs()), np.diff(self.agg_ptr))
        return out


def geometric_partition(dims: Sequence[int], block: Sequence[int] = (2, 2, 2)) -> Tuple[Partition, Tuple[int, int, int]]:
    """bx x by x bz boxes of a lexicographic grid (i = x + nx*(y + ny*z)); a trailing partial box
    joins its predecessor.  Returns the partition and the coarse grid dimensions."""
    from ._ffi import call, i64p, u64p

    nx, ny, nz = (int(d) for d in dims)
    bx, by, bz = (int(b) for b in block)
    coarse = np.zeros(3, dtype=np.int64)
    call("famg_geometric_partition", nx, ny, nz, bx, by, bz, None, None, coarse.ctypes.data_as(i64p))
    ptr = np.empty(int(coarse.prod()) + 1, dtype=np.uint64)
    nodes = np.empty(nx * ny * nz, dtype=np.uint64)
    call("famg_geometric_partition", nx, ny, nz, bx, by, bz, ptr.ctypes.data_as(u64p), nodes.ctypes.data_as(u64p),
         coarse.ctypes.data_as(i64p))
    # same bits as int64 (all values < 2^63): reinterpret, no copy
    return Partition(ptr.view(np.int64), nodes.view(np.int64), nx * ny * nz, validate=False), tuple(int(c) for c in coarse)


class DevicePartition:
    """Aggregate lists resident on the device (``famg_partition``): generated in place for box aggregates, or uploaded
    from any host :class:`Partition`.  The tentative prolongator is built from it on the device."""

    def __init__(self, ctx, handle):
        self._h, self.ctx = handle, ctx
        from ._ffi import call
        nn, na = C.c_int64(), C.c_int64()
        call("famg_partition_dims", handle, C.byref(nn), C.byref(na))
        self._nnodes, self._naggs = nn.value, na.value
        self._host = None

    @classmethod
    def geometric(cls, ctx, dims: Sequence[int], block: Sequence[int] = (2, 2, 2)):
        from ._ffi import call, i64p, vp
        h = vp()
        coarse = np.zeros(3, dtype=np.int64)
        call("famg_partition_geometric_dev", ctx._h, int(dims[0]), int(dims[1]), int(dims[2]), int(block[0]), int(block[1]), int(block[2]),
             C.byref(h), coarse.ctypes.data_as(i64p))
        return cls(ctx, h), tuple(int(c) for c in coarse)

    @classmethod
    def from_host(cls, ctx, part: Partition) -> "DevicePartition":
        from ._ffi import call, u64p, vp
        h = vp()
        ap = np.ascontiguousarray(part.agg_ptr, dtype=np.int64)
        an = np.ascontiguousarray(part.agg_nodes, dtype=np.int64)
        call("famg_partition_upload", ctx._h, part.nnodes(), part.naggs(), ap.ctypes.data_as(u64p), an.ctypes.data_as(u64p), C.byref(h))
        out = cls(ctx, h)
        out._host = part
        return out

    def naggs(self) -> int:
        return self._naggs

    def nnodes(self) -> int:
        return self._nnodes

    def to_host(self) -> Partition:
        if self._host is None:
            from ._ffi import call, u64p
            ptr = np.empty(self._naggs + 1, dtype=np.uint64)
            nodes = np.empty(max(self._nnodes, 1), dtype=np.uint64)
            call("famg_partition_download", self._h, ptr.ctypes.data_as(u64p), nodes.ctypes.data_as(u64p))
            self._host = Partition(ptr.view(np.int64), nodes.view(np.int64)[: self._nnodes], self._nnodes, validate=False)
        return self._host

    # the hot path only needs the lists; everything else goes through the host view
    def __getattr__(self, name):
        if name in ("agg_ptr", "agg_nodes", "aggregates", "node_to_agg", "validate"):
            return getattr(self.to_host(), name)
        raise AttributeError(name)

    def __del__(self):
        try:
            from . import _ffi
            _ffi.lib().famg_partition_destroy(self._h)
        except Exception:
            pass


class GeometricPartitioner:
    """Callable ``(level, op, near_null) -> Partition`` for :class:`HierarchyConfig`.  ``device(level, ctx)`` yields the
    same aggregates as a :class:`DevicePartition` generated on the device (used by the hierarchy build for scalar
    problems with one near-null vector)."""

    def __init__(self, dims: Sequence[int], block: Sequence[int] = (2, 2, 2)):
        self.dims = [tuple(dims)]
        self.block = tuple(block)

    def _coarse(self, d):
        return tuple(max(d[i] // self.block[i], 1) for i in range(3))

    def __call__(self, level: int, op, near_null) -> Partition:
        part, coarse = geometric_partition(self.dims[level], self.block)
        if len(self.dims) == level + 1:
            self.dims.append(coarse)
        return part

    def device(self, level: int, ctx) -> DevicePartition:
        part, coarse = DevicePartition.geometric(ctx, self.dims[level], self.block)
        if len(self.dims) == level + 1:
            self.dims.append(coarse)
        return part


class StrengthGraph:
    """``AdjacencyList`` (partitioners/mod.rs:331-334), host-resident."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def new_ls_strength_graph(cls, mat, near_null, weights, max_depth: int = 3) -> "StrengthGraph":
        """partitioners/mod.rs:337-393.  ``mat``: device ``SparseRowMat`` (pattern downloaded) or a
        ``(row_ptr, col_idx)`` pair of host arrays."""
        from ._ffi import call, f64p, u64p, vp
        from .core import as_colmajor

        if isinstance(mat, tuple):
            rp, ci = mat
        else:
            rp, ci, _ = mat.to_host()
        rp = np.ascontiguousarray(rp, dtype=np.uint64)
        ci = np.ascontiguousarray(ci, dtype=np.uint64)
        nn = as_colmajor(near_null)
        n = len(rp) - 1
        if nn.shape[0] != n:
            raise ValueError("near_null rows must match the matrix")  # mod.rs:284
        w = np.ascontiguousarray(weights, dtype=np.float64)
        if len(w) < nn.shape[1]:
            raise ValueError("one weight per near-null vector is required")
        h = vp()
        call("famg_strength_graph_create", n, rp.ctypes.data_as(u64p), ci.ctypes.data_as(u64p), nn.ctypes.data_as(f64p),
             max(nn.shape[0], 1), nn.shape[1], w.ctypes.data_as(f64p), max_depth, C.byref(h))
        return cls(h)

    @classmethod
    def from_csr(cls, row_ptr, col_idx, w) -> "StrengthGraph":
        from ._ffi import call, f64p, u64p, vp

        rp = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        ci = np.ascontiguousarray(col_idx, dtype=np.uint64)
        w = np.ascontiguousarray(w, dtype=np.float64)
        h = vp()
        call("famg_graph_create", len(rp) - 1, rp.ctypes.data_as(u64p), ci.ctypes.data_as(u64p), w.ctypes.data_as(f64p), C.byref(h))
        return cls(h)

    def block_reduce(self, block_size: int):
        """``strength.aggregate(&block_reduce); strength.filter_diag()`` (partitioners/mod.rs:293-300), in place."""
        from ._ffi import call

        call("famg_graph_block_reduce", self._h, int(block_size))

    def dims(self) -> Tuple[int, int]:
        from ._ffi import call

        n, nnz = C.c_int64(), C.c_int64()
        call("famg_graph_dims", self._h, C.byref(n), C.byref(nnz))
        return n.value, nnz.value

    def to_csr(self):
        """(row_ptr, neighbour ids, weights) -- neighbour lists ascending by id."""
        from ._ffi import call, f64p, u64p

        n, nnz = self.dims()
        rp = np.empty(n + 1, dtype=np.uint64)
        ci = np.empty(max(nnz, 1), dtype=np.uint64)
        w = np.empty(max(nnz, 1))
        call("famg_graph_download", self._h, rp.ctypes.data_as(u64p), ci.ctypes.data_as(u64p), w.ctypes.data_as(f64p))
        return rp.astype(np.int64), ci[:nnz].astype(np.int64), w[:nnz]

    def __del__(self):
        try:
            from . import _ffi

            _ffi.lib().famg_graph_destroy(self._h)
        except Exception:
            pass


class PartitionerConfig:
    """partitioners/mod.rs:249-329 (``callback`` is a visualisation hook and is not mirrored).
    Also usable as the ``partitioner`` callable of :class:`AggregationConfig`; there the
    coarsening factor is scaled by ``candidate_dimension / block_size`` exactly as
    ``AggregationConfig::build`` does (interpolation/mod.rs:135-137)."""

    def __init__(self, coarsening_factor: float = 8.0, agg_size_penalty: float = 1.0, max_improvement_iters: int = 100):
        self.coarsening_factor = float(coarsening_factor)
        self.agg_size_penalty = float(agg_size_penalty)
        self.max_improvement_iters = int(max_improvement_iters)

    def build_from_strength(self, strength: StrengthGraph) -> Partition:
        from ._ffi import call, u64p

        n, _ = strength.dims()
        node_to_agg = np.zeros(n, dtype=np.uint64)
        naggs = C.c_int64()
        call("famg_partition_modularity", strength._h, self.coarsening_factor, self.agg_size_penalty, self.max_improvement_iters,
             node_to_agg.ctypes.data_as(u64p), C.byref(naggs))
        part = Partition.from_node_to_agg(node_to_agg.astype(np.int64))
        assert part.naggs() == naggs.value
        return part

    def build_partition(self, mat, near_null, weights) -> Partition:
        """partitioners/mod.rs:319-328.  ``mat``: ``SparseMatOp`` (block size 1) or ``SparseRowMat``."""
        block_size = mat.block_size() if hasattr(mat, "block_size") else 1
        m = mat.mat_ref() if hasattr(mat, "mat_ref") else mat
        if m.nrows != m.ncols:
            raise ValueError("square matrix expected")  # mod.rs:283
        strength = StrengthGraph.new_ls_strength_graph(m, near_null, weights, 3)
        if block_size > 1:  # mod.rs:293-300: the partition is over nodes of block_size dofs
            strength.block_reduce(block_size)
        return self.build_from_strength(strength)

    def scaled(self, ratio: float) -> "PartitionerConfig":
        return PartitionerConfig(self.coarsening_factor * ratio, self.agg_size_penalty, self.max_improvement_iters)
