"""Multi-GPU: one process per GPU, 1-D row partition.

``torch.distributed`` is used for plumbing only (rendezvous, broadcasting the NCCL unique id); the
data path of the solve is libfamg's peer-memory (CUDA IPC over NVLink) halo exchange with NCCL as
bootstrap / fallback (``csrc/dist.cu``), the hierarchy is built on row slabs with NCCL exchanges of
the off-rank rows (``csrc/dist_setup.cu``): :class:`DistHierarchy` is ``Hierarchy::coarsen``
(hierarchy.rs:190-248) where every rank computes only its rows of P, R and A_c.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import CgInfoStruct, call, f64p, i64p, vp
from .core import Context, DeviceMat, SparseMatOp, SparseRowMat, _f
from .partitioners import GeometricPartitioner, Partition, geometric_partition
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
    def sim(cls, ctx: Context, nranks: int) -> "Comm":
        """All ``nranks`` virtual ranks hosted by this process on one GPU: runs the distributed hierarchy
        construction with device copies in place of the exchanges (tests on a single GPU; setup only)."""
        self = cls.__new__(cls)
        h = vp()
        call("famg_comm_create_sim", ctx._h, nranks, C.byref(h))
        self._h, self.ctx, self.nranks, self.rank = h, ctx, nranks, 0
        return self

    @property
    def nlocal(self) -> int:
        n = C.c_int()
        call("famg_comm_dims", self._h, None, None, C.byref(n))
        return n.value

    def vranks(self) -> List[int]:
        """Virtual ranks hosted by this process."""
        return list(range(self.nranks)) if self.nlocal > 1 else [self.rank]

    def allgatherv(self, pieces: Sequence[np.ndarray]) -> np.ndarray:
        """Concatenation over all ranks (rank order) of one float64 array per hosted rank."""
        nl = self.nlocal
        loc = [np.ascontiguousarray(p, dtype=np.float64).reshape(-1) for p in pieces]
        assert len(loc) == nl
        counts = np.asarray([len(p) for p in loc], dtype=np.int64)
        # total length: sum over all ranks; the ABI fills one full copy per hosted rank
        import_total = self._allgather_counts(counts)
        outs = [np.empty(max(import_total, 1)) for _ in range(nl)]
        lp = (f64p * nl)(*[p.ctypes.data_as(f64p) for p in loc])
        op = (f64p * nl)(*[o.ctypes.data_as(f64p) for o in outs])
        call("famg_comm_allgatherv_f64", self._h, lp, counts.ctypes.data_as(i64p), op)
        return outs[0][:import_total]

    def _allgather_counts(self, counts: np.ndarray) -> int:
        if self.nlocal > 1 or self.nranks == 1:
            return int(counts.sum())
        return int(round(self.allreduce_sum([float(counts[0])])[0]))

    @classmethod
    def from_torch(cls, ctx: Context) -> "Comm":
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return cls(ctx, 1, 0, None)
        # torchrun exports OMP_NUM_THREADS=1: give every rank its share of the host cores for the host-side setup pieces
        import os
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        local = int(os.environ.get("LOCAL_WORLD_SIZE", dist.get_world_size()))
        call("famg_set_num_threads", max(1, cores // max(local, 1)))
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

    @classmethod
    def from_hierarchy(cls, comm: Comm, h: "DistHierarchy", smoother: str = "l1", omega: float = 0.66, mu: int = 1,
                       smoothing_steps: int = 1) -> "DistMultigrid":
        """``MultigridConfig::build`` (multigrid.rs:52-164) over a :class:`DistHierarchy`: diagonal smoothers
        on the distributed levels, the replicated tail as an ordinary :class:`Multigrid` (exact coarsest solve)."""
        from .preconditioners.multigrid import MultigridConfig
        kinds = {"l1": 0, "jacobi": 2}
        if smoother not in kinds:
            raise ValueError("distributed levels support the 'l1' and 'jacobi' smoothers")
        tail_mg = MultigridConfig(mu=mu, smoothing_steps=smoothing_steps, smoother=smoother, omega=omega).build(h.tail, is_tail=True)
        n = h.n_dist
        a = (vp * n)(*[m._h for m in h.A])
        r = (vp * n)(*[m._h for m in h.R])
        p = (vp * n)(*[m._h for m in h.P])
        self = cls.__new__(cls)
        hd = vp()
        call("famg_dist_mg_create_levels", comm._h, n, a, r, p, kinds[smoother], float(omega), tail_mg._h, C.byref(hd))
        self._h, self.comm, self.global_mg, self.hierarchy = hd, comm, tail_mg, h
        rs = h.A[0].row_split()
        self._splits = [rs]
        self.row_begin, self.row_end = int(rs[comm.rank]), int(rs[comm.rank + 1])
        return self

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


class DistMat:
    """Row-partitioned sparse matrix (``famg_dmat``): one slab per hosted virtual rank."""

    def __init__(self, comm: Comm, handle):
        self._h, self.comm = handle, comm

    @classmethod
    def from_slabs(cls, comm: Comm, slabs: Sequence[SparseRowMat], ncols_global: int, col_split=None) -> "DistMat":
        """Slabs carry GLOBAL column ids and are taken over (their columns are renumbered in place by
        :meth:`finalize`)."""
        arr = (vp * len(slabs))(*[s._h for s in slabs])
        cs = None if col_split is None else np.ascontiguousarray(col_split, dtype=np.int64)
        h = vp()
        call("famg_dmat_create", comm._h, arr, int(ncols_global), None if cs is None else cs.ctypes.data_as(i64p), C.byref(h))
        return cls(comm, h)

    def finalize(self, replicated_cols: bool = False) -> "DistMat":
        call("famg_dmat_finalize", self._h, 1 if replicated_cols else 0)
        return self

    def info(self):
        nr, nc = C.c_int64(), C.c_int64()
        rs = np.zeros(self.comm.nranks + 1, dtype=np.int64)
        cs = np.zeros(self.comm.nranks + 1, dtype=np.int64)
        call("famg_dmat_info", self._h, C.byref(nr), C.byref(nc), rs.ctypes.data_as(i64p), cs.ctypes.data_as(i64p))
        return nr.value, nc.value, rs, cs

    @property
    def nrows(self) -> int:
        return self.info()[0]

    @property
    def ncols(self) -> int:
        return self.info()[1]

    def row_split(self) -> np.ndarray:
        return self.info()[2]

    def local(self, li: int = 0, global_cols: bool = False) -> SparseRowMat:
        h = vp()
        call("famg_dmat_local", self._h, li, 1 if global_cols else 0, C.byref(h))
        return SparseRowMat(self.comm.ctx, h)

    def gather(self) -> SparseRowMat:
        """The whole matrix, replicated (collective)."""
        nl = self.comm.nlocal
        arr = (vp * nl)()
        call("famg_dmat_gather", self._h, arr)
        mats = [SparseRowMat(self.comm.ctx, vp(arr[i])) for i in range(nl)]
        return mats[0]

    def __del__(self):
        try:
            _ffi.lib().famg_dmat_destroy(self._h)
        except Exception:
            pass


class DistGeometricPartitioner:
    """Box aggregates of a lexicographic grid cut into z-slabs: every rank aggregates its own planes
    (boxes never straddle ranks) -- the same aggregates, in the same global order, as
    :class:`GeometricPartitioner` on the whole grid."""

    def __init__(self, dims: Sequence[int], block: Sequence[int] = (2, 2, 2)):
        self.dims = [tuple(int(d) for d in dims)]
        self.block = tuple(int(b) for b in block)

    def level_dims(self, level: int):
        while len(self.dims) <= level:
            d = self.dims[-1]
            self.dims.append(tuple(max(d[i] // self.block[i], 1) for i in range(3)))
        return self.dims[level]

    def local(self, level: int, row_begin: int, row_end: int) -> Partition:
        nx, ny, nz = self.level_dims(level)
        bz = self.block[2]
        plane = nx * ny
        if row_begin % plane or row_end % plane:
            raise ValueError("row slabs of a lexicographic grid must consist of whole z-planes")
        z0, z1 = row_begin // plane, row_end // plane
        cz = max(nz // bz, 1)
        ok = z0 % bz == 0 and (z1 == nz or (z1 % bz == 0 and z1 // bz < cz)) and z1 > z0
        if not ok:
            raise ValueError(f"slab planes [{z0}, {z1}) of {nz} do not hold whole {bz}-plane boxes: aggregates would straddle ranks")
        part, _ = geometric_partition((nx, ny, z1 - z0), self.block)
        return part

    def local_dev(self, ctx, level: int, row_begin: int, row_end: int):
        """The same aggregates generated on the device (:class:`DevicePartition`)."""
        from .partitioners import DevicePartition
        nx, ny, nz = self.level_dims(level)
        bz = self.block[2]
        plane = nx * ny
        if row_begin % plane or row_end % plane:
            raise ValueError("row slabs of a lexicographic grid must consist of whole z-planes")
        z0, z1 = row_begin // plane, row_end // plane
        cz = max(nz // bz, 1)
        if not (z0 % bz == 0 and (z1 == nz or (z1 % bz == 0 and z1 // bz < cz)) and z1 > z0):
            raise ValueError(f"slab planes [{z0}, {z1}) of {nz} do not hold whole {bz}-plane boxes: aggregates would straddle ranks")
        return DevicePartition.geometric(ctx, (nx, ny, z1 - z0), self.block)[0]

    def tail(self, level: int) -> GeometricPartitioner:
        """Partitioner of the replicated levels, whose level 0 is this hierarchy's ``level``."""
        return GeometricPartitioner(self.level_dims(level), self.block)


def fine_plane_splits(dims: Sequence[int], nranks: int, block_z: int = 2, max_levels: int = 32) -> np.ndarray:
    """Row split of the fine grid into z-slabs whose cuts stay box-aligned on as many levels as possible
    (same rule as :func:`level_row_splits`)."""
    nz0 = dims[2]
    gran = 1
    for lvl in range(1, max_levels):
        g = block_z ** lvl
        if nz0 % g == 0 and nz0 // g >= nranks:
            gran = g
    return slab_splits(dims, nranks, gran)


class DistHierarchy:
    """``Hierarchy`` (hierarchy.rs:61-360) on row slabs.  Levels ``0 .. n_dist-1`` are distributed
    (:class:`DistMat` operators / transfer operators, near-null slices per hosted rank); from level
    ``n_dist`` on the hierarchy is replicated (``tail``: an ordinary :class:`Hierarchy` every rank builds
    from the gathered operator -- a few thousand rows).  Scalar problems, one near-null vector."""

    def __init__(self, comm: Comm, a0: DistMat, near_null_local: Sequence[np.ndarray], partitioner: DistGeometricPartitioner,
                 coarsest_dim: int = 1000, replicate_below: int = 4096, smoothing_steps: int = 1, jacobi_weight: float = 0.66):
        self.comm = comm
        self.A: List[DistMat] = [a0]
        self.P: List[DistMat] = []
        self.R: List[DistMat] = []
        self.near_nulls: List[list] = [[np.ascontiguousarray(v, dtype=np.float64).reshape(-1) for v in near_null_local]]
        # aggregates and tentative prolongator on the device when the partitioner can (FAMG_HOST_AGGREGATES=1: host path)
        import os
        self.device_aggregates = hasattr(partitioner, "local_dev") and os.environ.get("FAMG_HOST_AGGREGATES") is None
        self.partitioner = partitioner
        self.coarsest_dim, self.replicate_below = coarsest_dim, replicate_below
        self.smoothing_steps, self.jacobi_weight = smoothing_steps, jacobi_weight
        self.tail = None  # replicated Hierarchy of the levels >= n_dist
        if not a0.info()[0] // comm.nranks >= replicate_below:
            raise ValueError("problem too small to partition: use Hierarchy on one GPU")
        a0.finalize(False)
        self.coarsen()

    @property
    def n_dist(self) -> int:
        return len(self.A)

    def levels(self) -> int:
        return self.n_dist + (self.tail.levels() if self.tail is not None else 0)

    def near_null_host(self, level: int) -> np.ndarray:
        """This process' near-null slices of a distributed level, concatenated (host copy)."""
        return np.concatenate([v.to_host().ravel() if hasattr(v, "to_host") else np.asarray(v).ravel() for v in self.near_nulls[level]])

    def _coarsen_once_dev(self, level: int):
        comm, nl = self.comm, self.comm.nlocal
        fine = self.A[level]
        rs = fine.row_split()
        ctx = comm.ctx
        parts = [self.partitioner.local_dev(ctx, level, int(rs[r]), int(rs[r + 1])) for r in comm.vranks()]
        nn = [v if hasattr(v, "_h") else DeviceMat.from_host(ctx, v) for v in self.near_nulls[level]]
        cnn = [DeviceMat(ctx, p.naggs(), 1) for p in parts]
        pa = (vp * nl)(*[p._h for p in parts])
        nv = (vp * nl)(*[v._h for v in nn])
        cv = (vp * nl)(*[v._h for v in cnn])
        p, r, ac = vp(), vp(), vp()
        call("famg_dist_coarsen_dev", fine._h, pa, nv, self.smoothing_steps, float(self.jacobi_weight), C.byref(p), C.byref(r), C.byref(ac), cv)
        return DistMat(comm, p), DistMat(comm, r), DistMat(comm, ac), cnn

    def _coarsen_once(self, level: int):
        if self.device_aggregates:
            return self._coarsen_once_dev(level)
        comm, nl = self.comm, self.comm.nlocal
        fine = self.A[level]
        rs = fine.row_split()
        parts = [self.partitioner.local(level, int(rs[r]), int(rs[r + 1])) for r in comm.vranks()]
        nn = self.near_nulls[level]
        n_aggs = np.asarray([p.naggs() for p in parts], dtype=np.int64)
        ptrs = [np.ascontiguousarray(p.agg_ptr, dtype=np.int64) for p in parts]
        nodes = [np.ascontiguousarray(p.agg_nodes, dtype=np.int64) for p in parts]
        cnn = [np.zeros(max(int(n), 1)) for n in n_aggs]
        from ._ffi import u64p
        ap = (u64p * nl)(*[a.ctypes.data_as(u64p) for a in ptrs])
        an = (u64p * nl)(*[a.ctypes.data_as(u64p) for a in nodes])
        nv = (f64p * nl)(*[v.ctypes.data_as(f64p) for v in nn])
        cv = (f64p * nl)(*[v.ctypes.data_as(f64p) for v in cnn])
        p, r, ac = vp(), vp(), vp()
        call("famg_dist_coarsen", fine._h, n_aggs.ctypes.data_as(i64p), ap, an, nv, self.smoothing_steps, float(self.jacobi_weight),
             C.byref(p), C.byref(r), C.byref(ac), cv)
        cnn = [v[: int(n)] for v, n in zip(cnn, n_aggs)]
        return DistMat(comm, p), DistMat(comm, r), DistMat(comm, ac), cnn

    def coarsen(self):
        """hierarchy.rs:190-248 with the level loop running on slabs while a level has at least
        ``replicate_below`` rows per rank."""
        from .hierarchy import Hierarchy, HierarchyConfig, thin_q
        from .interpolation import AggregationConfig
        from .preconditioners.smoothers import StationaryIteration, new_l1
        import os, time
        comm = self.comm
        trace = os.environ.get("FAMG_SETUP_TRACE") is not None and comm.rank == 0

        class T:
            def __init__(s, what): s.what = what
            def __enter__(s):
                if trace: comm.ctx.sync(); s.t = time.perf_counter()
            def __exit__(s, *a):
                if trace: comm.ctx.sync(); print(f"[setup] {s.what:36s} {1e3 * (time.perf_counter() - s.t):9.3f} ms", file=__import__("sys").stderr, flush=True)
        level = 0
        while True:
            with T(f"level {level}: coarsen (partition + dist_coarsen)"):
                P, R, Ac, cnn = self._coarsen_once(level)
            nc = Ac.info()[0]
            self.P.append(P); self.R.append(R)
            if nc // comm.nranks < self.replicate_below or nc <= self.coarsest_dim:
                # transition to the replicated tail: gather A_c and the coarse near-null, then the reference's own loop
                with T(f"level {level}: gather A_c + near-null"):
                    P.finalize(True)
                    g = Ac.gather()
                    nn_g = comm.allgatherv([v.to_host().ravel() if hasattr(v, "to_host") else v for v in cnn]).reshape(-1, 1)
                with T("replicated tail (reference loop)"):
                    l1 = new_l1(g)
                    nn_dev = DeviceMat.from_host(g.ctx, nn_g)
                    StationaryIteration(g, l1, 3).apply_in_place_dev(nn_dev)  # hierarchy.rs:217-226
                    nn_q = thin_q(nn_dev.to_host())                            # :228
                    cfg = HierarchyConfig(self.coarsest_dim, AggregationConfig(self.smoothing_steps, 1, self.partitioner.tail(level + 1)))
                    self.tail = Hierarchy(SparseMatOp(g), nn_q, None, cfg)
                    if nc > self.coarsest_dim:  # the level loop's own test (hierarchy.rs:199), already decided for this level
                        self.tail.coarsen()
                return
            with T(f"level {level}: halo plans of A_c and P"):
                Ac.finalize(False)
                P.finalize(False)
            nl = comm.nlocal
            with T(f"level {level}: near-null smoothing + thin Q"):
                if self.device_aggregates:
                    call("famg_dist_smooth_near_null_dev", Ac._h, 3, (vp * nl)(*[v._h for v in cnn]))
                else:
                    arr = (f64p * nl)(*[np.ascontiguousarray(v).ctypes.data_as(f64p) for v in cnn])
                    call("famg_dist_smooth_near_null", Ac._h, 3, arr)
            self.A.append(Ac)
            self.near_nulls.append(cnn)
            level += 1
