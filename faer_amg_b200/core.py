"""Operators of the hot path: ``SparseMatOp`` / ``ParSpmmOp`` / device matrices.

Mirrors ``src/core.rs`` and ``src/par_spmm.rs`` of the reference.  ``SparseRowMat`` is the device
CSR (faer ``SparseRowMat<usize, f64>``), ``DeviceMat`` a device-resident ``Mat<f64>``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _ffi
from ._ffi import call, f64p, u64p, vp

PAR_BLOCK_SIZE = 8192  # par_spmm.rs:15 (kept for API parity; the device kernel tiles by CTA)


def _f(a):
    return a.ctypes.data_as(f64p)


def _u(a):
    return a.ctypes.data_as(u64p)


def as_colmajor(x) -> np.ndarray:
    """n x k float64, column-major, unit row stride (faer Mat layout)."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    return np.asfortranarray(x)


class Context:
    """One per (process, device): the CUDA stream all work is ordered on.  Takes the place of
    ``faer::set_global_parallelism(Par::Rayon(n))`` (examples/amg/main.rs:227-228)."""

    _default: dict = {}

    def __init__(self, device: int = 0):
        h = vp()
        call("famg_ctx_create", device, C.byref(h))
        self._h = h
        self.device = device

    @classmethod
    def default(cls, device: int = 0) -> "Context":
        if device not in cls._default:
            cls._default[device] = cls(device)
        return cls._default[device]

    def sync(self):
        call("famg_ctx_sync", self._h)

    @property
    def stream(self) -> int:
        s = vp()
        call("famg_ctx_stream", self._h, C.byref(s))
        return s.value or 0

    def info(self) -> dict:
        sms, free, total = C.c_int(), C.c_int64(), C.c_int64()
        name = C.create_string_buffer(256)
        call("famg_ctx_info", self._h, C.byref(sms), C.byref(free), C.byref(total), name, 256)
        return {"num_sms": sms.value, "mem_free": free.value, "mem_total": total.value, "name": name.value.decode()}

    def set_option(self, key: str, value: int):
        """Kernel-selection knobs for A/B measurements (see famg_ctx_set_option)."""
        call("famg_ctx_set_option", self._h, key.encode(), int(value))

    def reserve(self, nbytes: int):
        """Pre-size the device memory pool (see famg_ctx_reserve)."""
        call("famg_ctx_reserve", self._h, int(nbytes))

    def trace_dump(self, path: str):
        """Text dump of the in-kernel timeline collected since ``set_option("trace", 1)``."""
        call("famg_ctx_trace_dump", self._h, path.encode())

    def launch_count(self) -> int:
        n = C.c_int64()
        call("famg_ctx_launch_count", self._h, C.byref(n))
        return n.value


class DeviceMat:
    """Device-resident column-major ``Mat<f64>`` (n x k)."""

    def __init__(self, ctx: Context, nrows: int, ncols: int = 1):
        h = vp()
        call("famg_vec_create", ctx._h, nrows, ncols, C.byref(h))
        self._h, self.ctx, self.nrows, self.ncols = h, ctx, nrows, ncols

    @classmethod
    def from_host(cls, ctx: Context, x) -> "DeviceMat":
        x = as_colmajor(x)
        m = cls(ctx, x.shape[0], x.shape[1])
        m.upload(x)
        return m

    def upload(self, x):
        x = as_colmajor(x)
        assert x.shape == (self.nrows, self.ncols)
        call("famg_vec_upload", self._h, _f(x), max(x.shape[0], 1))

    def to_host(self) -> np.ndarray:
        out = np.empty((self.nrows, self.ncols), order="F")
        call("famg_vec_download", self._h, _f(out), max(self.nrows, 1))
        return out

    def fill(self, v: float):
        call("famg_vec_fill", self._h, float(v))

    def copy_from(self, other: "DeviceMat"):
        call("famg_vec_copy", self._h, other._h)

    def axpby(self, alpha: float, x: "DeviceMat", beta: float):
        """self = alpha * x + beta * self."""
        call("famg_vec_axpby", self._h, float(alpha), x._h, float(beta))

    def norm_l2(self) -> np.ndarray:
        out = np.empty(self.ncols)
        call("famg_vec_norm2", self._h, _f(out))
        return out

    def data_ptr(self) -> Tuple[int, int]:
        p, ld = vp(), C.c_int64()
        call("famg_vec_ptr", self._h, C.byref(p), C.byref(ld))
        return (p.value or 0), ld.value

    def __del__(self):
        try:
            _ffi.lib().famg_vec_destroy(self._h)
        except Exception:
            pass


class SparseRowMat:
    """Device CSR.  Host-side constructors take the reference's ``usize`` indices."""

    def __init__(self, ctx: Context, handle):
        self._h, self.ctx = handle, ctx
        nr, nc, nnz = C.c_int64(), C.c_int64(), C.c_int64()
        call("famg_csr_dims", handle, C.byref(nr), C.byref(nc), C.byref(nnz))
        self.nrows, self.ncols, self.nnz = nr.value, nc.value, nnz.value

    @property
    def shape(self):
        return (self.nrows, self.ncols)

    def compute_nnz(self) -> int:
        return self.nnz

    @classmethod
    def from_csr(cls, ctx: Context, nrows, ncols, row_ptr, col_idx, val) -> "SparseRowMat":
        rp = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        ci = np.ascontiguousarray(col_idx, dtype=np.uint64)
        v = np.ascontiguousarray(val, dtype=np.float64)
        h = vp()
        call("famg_csr_create", ctx._h, nrows, ncols, _u(rp), _u(ci), _f(v), C.byref(h))
        return cls(ctx, h)

    @classmethod
    def try_new_from_triplets(cls, ctx: Context, nrows, ncols, rows, cols, vals) -> "SparseRowMat":
        r = np.ascontiguousarray(rows, dtype=np.uint64)
        c = np.ascontiguousarray(cols, dtype=np.uint64)
        v = np.ascontiguousarray(vals, dtype=np.float64)
        h = vp()
        call("famg_csr_create_from_triplets", ctx._h, nrows, ncols, len(r), _u(r), _u(c), _f(v), C.byref(h))
        return cls(ctx, h)

    @classmethod
    def from_scipy(cls, ctx: Context, m) -> "SparseRowMat":
        m = m.tocsr()
        m.sort_indices()
        return cls.from_csr(ctx, m.shape[0], m.shape[1], m.indptr, m.indices, m.data)

    def to_host(self):
        """(row_ptr, col_idx, val) with uint64 indices, caller-owned."""
        rp = np.empty(self.nrows + 1, dtype=np.uint64)
        ci = np.empty(max(self.nnz, 1), dtype=np.uint64)
        v = np.empty(max(self.nnz, 1))
        call("famg_csr_download", self._h, _u(rp), _u(ci), _f(v))
        return rp, ci[: self.nnz], v[: self.nnz]

    def to_scipy(self):
        import scipy.sparse as sp
        rp, ci, v = self.to_host()
        return sp.csr_matrix((v, ci.astype(np.int64), rp.astype(np.int64)), shape=self.shape)

    def plan(self) -> dict:
        tpr, rows, mx = C.c_int(), C.c_int(), C.c_int()
        avg = C.c_double()
        call("famg_csr_plan", self._h, C.byref(tpr), C.byref(rows), C.byref(avg), C.byref(mx))
        return {"threads_per_row": tpr.value, "rows_per_cta": rows.value, "avg_row_nnz": avg.value, "max_row_nnz": mx.value}

    def row_slab(self, r0: int, r1: int) -> "SparseRowMat":
        h = vp()
        call("famg_csr_row_slab", self._h, r0, r1, C.byref(h))
        return SparseRowMat(self.ctx, h)

    def transpose(self) -> "SparseRowMat":
        """``p.transpose().to_row_major()`` (interpolation/mod.rs:824-827)."""
        h = vp()
        call("famg_transpose", self._h, C.byref(h))
        return SparseRowMat(self.ctx, h)

    def __matmul__(self, other):
        """``&a * &b`` for sparse operands (interpolation/mod.rs:828); dense operands -> SpMM."""
        if isinstance(other, SparseRowMat):
            h = vp()
            call("famg_spgemm", self._h, other._h, C.byref(h))
            return SparseRowMat(self.ctx, h)
        return self.apply(other)

    # LinOp surface -----------------------------------------------------------------------
    def apply(self, rhs) -> np.ndarray:
        """``LinOp::apply(out, rhs)`` with host buffers (staged through the device)."""
        rhs = as_colmajor(rhs)
        out = np.empty((self.nrows, rhs.shape[1]), order="F")
        call("famg_spmm", self._h, _f(out), max(self.nrows, 1), _f(rhs), max(rhs.shape[0], 1), rhs.shape[1])
        return out

    def apply_dev(self, out: DeviceMat, rhs: DeviceMat):
        call("famg_spmm_dev", self._h, out._h, rhs._h)

    def residual_dev(self, out: DeviceMat, b: DeviceMat, x: DeviceMat):
        call("famg_residual_dev", self._h, out._h, b._h, x._h)

    def apply_add_dev(self, y: DeviceMat, x: DeviceMat):
        call("famg_spmm_add_dev", self._h, y._h, x._h)

    def time_kernel(self, which: int, reps: int = 50, warmup: int = 5) -> float:
        """Average milliseconds per launch of kernel class 0 SpMV / 1 residual / 2 smoother sweep,
        CUDA events on the context stream."""
        ms = C.c_float()
        call("famg_time_kernel", self._h, which, reps, warmup, C.byref(ms))
        return ms.value

    def __del__(self):
        try:
            _ffi.lib().famg_csr_destroy(self._h)
        except Exception:
            pass


class ParSpmmOp:
    """``ParSpmmOp`` (par_spmm.rs:17-159): the accelerated mat-apply ``LinOp``.  On the B200 the
    tiling into 8192-row CSC blocks is replaced by the CTA-tiled CSR kernel of ``csrc/spmv.cu``;
    the object is a view of the same device CSR (no second copy of the matrix, unlike core.rs:38-41).
    Works for rectangular operators (the reference does not, F10a)."""

    def __init__(self, mat: SparseRowMat):
        self.mat = mat

    def nrows(self) -> int:
        return self.mat.nrows

    def ncols(self) -> int:
        return self.mat.ncols

    def apply(self, rhs) -> np.ndarray:
        return self.mat.apply(rhs)

    conj_apply = apply

    def apply_dev(self, out: DeviceMat, rhs: DeviceMat):
        self.mat.apply_dev(out, rhs)


class SparseMatOp:
    """``SparseMatOp`` (core.rs:13-111): square CSR + block size + accelerated op."""

    def __init__(self, mat: SparseRowMat, block_size: int = 1):
        if mat.nrows != mat.ncols:  # core.rs:57-59
            raise ValueError(
                f"SparseMatOp is only designed for square sparse matrices. Matrix dimensions are {mat.nrows}x{mat.ncols}")
        self._check_block_size(mat, block_size)
        self.mat = mat
        self._block_size = block_size
        self._par_op = ParSpmmOp(mat)

    @staticmethod
    def _check_block_size(mat, block_size):  # core.rs:103-110
        if mat.nrows % block_size != 0:
            raise ValueError(
                f"Matrix is incompatible with provided block size. `mat.nrows() % block_size = {mat.nrows % block_size}` and should be 0.")

    def mat_ref(self) -> SparseRowMat:
        return self.mat

    arc_mat = mat_ref

    def par_op(self) -> Optional[ParSpmmOp]:
        return self._par_op

    def dyn_op(self) -> ParSpmmOp:
        return self._par_op

    def block_size(self) -> int:
        return self._block_size

    def set_block_size(self, block_size: int):
        self._check_block_size(self.mat, block_size)
        self._block_size = block_size
