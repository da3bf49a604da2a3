"""Multi-GPU: one process per GPU, 1-D row partition, NCCL halo exchange.

``torch.distributed`` is used for plumbing only (rendezvous, broadcasting the NCCL unique id,
host-side gathers); the data path is libfamg's own NCCL communicator (``csrc/dist.cu``).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import CgInfoStruct, call, f64p, i64p, vp
from .core import Context, DeviceMat, _f
from .solvers import CgInfo, CgParams, _finish


def slab_splits(dims: Sequence[int], nranks: int, z_granule: int = 1) -> np.ndarray:
    """Row ranges of a lexicographic nx*ny*nz grid cut into z-slabs: rank r owns planes
    [z_r, z_{r+1}) with every cut a multiple of ``z_granule`` planes."""
    nx, ny, nz = dims
    units = nz // z_granule
    cuts = [(units * r) // nranks * z_granule for r in range(nranks)] + [nz]
    return np.asarray([c * nx * ny for c in cuts], dtype=np.int64)


def level_row_splits(level_dims: Sequence[Sequence[int]], nranks: int, block_z: int = 2) -> List[np.ndarray]:
    """Conformal splits for a geometric hierarchy: the fine cuts fall on multiples of
    block_z^(levels-1) planes where possible, so aggregates never straddle ranks and level l+1's
    owned rows are exactly the aggregates of level l's owned rows."""
    out = []
    nz0 = level_dims[0][2]
    # cut positions in fine planes, aligned to the coarsest granularity that still gives every rank work
    gran = 1
    for lvl in range(1, len(level_dims)):
        g = block_z ** lvl
        if nz0 % g == 0 and nz0 // g >= nranks:
            gran = g
    fine = slab_splits(level_dims[0], nranks, gran)
    planes = fine // (level_dims[0][0] * level_dims[0][1])
    for lvl, d in enumerate(level_dims):
        scale = block_z ** lvl
        z = np.minimum(planes // scale, d[2])
        z[-1] = d[2]
        z = np.maximum.accumulate(z)
        out.append((z * d[0] * d[1]).astype(np.int64))
    return out


def broadcast_unique_id(make: Optional[Callable[[], bytes]] = None) -> bytes:
    """Rank 0 creates the 128-byte NCCL id, everyone receives it over torch.distributed."""
    import torch
    import torch.distributed as dist
    if make is None:
        def make():
            buf = C.create_string_buffer(128)
            call("famg_comm_unique_id", buf)
            return buf.raw
    t = torch.zeros(128, dtype=torch.uint8)
    if dist.get_rank() == 0:
        t = torch.tensor(list(make()), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0)
    return bytes(t.cpu().tolist())


def gather_rows(local: np.ndarray, splits: np.ndarray) -> np.ndarray:
    """Host-side all-gather of row-partitioned data (setup / verification only)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    pieces: List = [None] * world
    dist.all_gather_object(pieces, np.asarray(local))
    return np.concatenate(pieces)


class Comm:
    """One rank of the job: libfamg's NCCL communicator on this rank's context."""

    def __init__(self, ctx: Context, nranks: int, rank: int, unique_id: Optional[bytes]):
        h = vp()
        buf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        call("famg_comm_create", ctx._h, nranks, rank, buf, C.byref(h))
        self._h, self.ctx, self.nranks, self.rank = h, ctx, nranks, rank

    @classmethod
    def from_torch(cls, ctx: Context) -> "Comm":
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return cls(ctx, 1, 0, None)
        return cls(ctx, dist.get_world_size(), dist.get_rank(), broadcast_unique_id())

    def allreduce_sum(self, vals) -> np.ndarray:
        v = np.ascontiguousarray(vals, dtype=np.float64).copy()
        call("famg_comm_allreduce_sum", self._h, _f(v), len(v))
        return v

    def __del__(self):
        try:
            _ffi.lib().famg_comm_destroy(self._h)
        except Exception:
            pass


class DistMultigrid:
    """Row-partitioned Multigrid + PCG built from a replicated global :class:`Multigrid`."""

    def __init__(self, comm: Comm, global_mg, row_splits: Sequence[np.ndarray], replicate_below: int = 4096):
        self._splits = [np.ascontiguousarray(s, dtype=np.int64) for s in row_splits]
        arr = (i64p * len(self._splits))(*[s.ctypes.data_as(i64p) for s in self._splits])
        h = vp()
        call("famg_dist_mg_create", comm._h, global_mg._h, arr, int(replicate_below), C.byref(h))
        self._h, self.comm, self.global_mg = h, comm, global_mg
        r = comm.rank
        self.row_begin, self.row_end = int(self._splits[0][r]), int(self._splits[0][r + 1])

    @property
    def nloc(self) -> int:
        return self.row_end - self.row_begin

    def spmv_dev(self, y: DeviceMat, x: DeviceMat):
        call("famg_dist_spmv_dev", self._h, y._h, x._h)

    def apply_dev(self, out: DeviceMat, rhs: DeviceMat):
        call("famg_dist_mg_apply_dev", self._h, out._h, rhs._h)

    def solve(self, x_local: np.ndarray, b_local, params: CgParams) -> CgInfo:
        b = np.ascontiguousarray(np.asarray(b_local, dtype=np.float64).reshape(-1))
        info = CgInfoStruct()
        st = _ffi.lib().famg_dist_pcg_solve(self._h, _f(x_local.reshape(-1)), _f(b), params.rel_tolerance, params.abs_tolerance,
                                             params.max_iters, 1 if params.initial_guess_zero else 0, C.byref(info))
        return _finish(st, info, params.max_iters)

    def solve_dev(self, x: DeviceMat, b: DeviceMat, params: CgParams) -> CgInfo:
        info = CgInfoStruct()
        st = _ffi.lib().famg_dist_pcg_solve_dev(self._h, x._h, b._h, params.rel_tolerance, params.abs_tolerance,
                                                 params.max_iters, 1 if params.initial_guess_zero else 0, C.byref(info))
        return _finish(st, info, params.max_iters)

    def __del__(self):
        try:
            _ffi.lib().famg_dist_mg_destroy(self._h)
        except Exception:
            pass
