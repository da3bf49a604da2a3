"""``Partition`` container + deterministic geometric aggregation.

The reference's algebraic partitioner (``src/partitioners/``) is host-side graph work that the
north-star keeps out of the GPU path; it is also non-deterministic.  The hot path only
consumes its *output*, a ``Partition`` (``partitioners/mod.rs:23-27``: node_to_agg + agg_to_node),
so that is what this module provides, plus the structured-grid aggregates used by the benchmark
configurations.
"""
from __future__ import annotations

from typing import Sequence, Tuple

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
    nx, ny, nz = dims
    bx, by, bz = block
    cx, cy, cz = max(nx // bx, 1), max(ny // by, 1), max(nz // bz, 1)
    if nx == cx * bx and ny == cy * by and nz == cz * bz:
        # regular case, built directly (no sort): aggregate (X,Y,Z) lists its nodes in (dz,dy,dx)
        # order, which is ascending node order
        ix = (bx * np.arange(cx, dtype=np.int64)[:, None] + np.arange(bx, dtype=np.int64)[None, :])
        iy = (by * np.arange(cy, dtype=np.int64)[:, None] + np.arange(by, dtype=np.int64)[None, :]) * nx
        iz = (bz * np.arange(cz, dtype=np.int64)[:, None] + np.arange(bz, dtype=np.int64)[None, :]) * (nx * ny)
        nodes = (iz[:, None, None, :, None, None] + iy[None, :, None, None, :, None] + ix[None, None, :, None, None, :])
        ptr = np.arange(cx * cy * cz + 1, dtype=np.int64) * (bx * by * bz)
        return Partition(ptr, nodes.reshape(-1), nx * ny * nz, validate=False), (cx, cy, cz)
    x = np.minimum(np.arange(nx) // bx, cx - 1)
    y = np.minimum(np.arange(ny) // by, cy - 1)
    z = np.minimum(np.arange(nz) // bz, cz - 1)
    agg = (x[None, None, :] + cx * (y[None, :, None] + cy * z[:, None, None])).reshape(-1)
    return Partition.from_node_to_agg(agg), (cx, cy, cz)


class GeometricPartitioner:
    """Callable ``(level, op, near_null) -> Partition`` for :class:`HierarchyConfig`."""

    def __init__(self, dims: Sequence[int], block: Sequence[int] = (2, 2, 2)):
        self.dims = [tuple(dims)]
        self.block = tuple(block)

    def __call__(self, level: int, op, near_null) -> Partition:
        part, coarse = geometric_partition(self.dims[level], self.block)
        if len(self.dims) == level + 1:
            self.dims.append(coarse)
        return part
