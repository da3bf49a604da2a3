"""CPU oracle for the faer-amg hot path -- TEST INFRASTRUCTURE ONLY.

ctypes wrapper around ``oracle/famg_oracle.c`` (the C restatement of the reference's algorithms;
see that file's header: *parity unpinned* -- the Rust reference cannot be built here and ships no
tests or golden vectors).  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package.  Nothing under
``faer_amg_b200/`` does.

The hierarchy driver (:func:`build_hierarchy`) restates ``Hierarchy::coarsen``
(``src/hierarchy.rs:190-248``) and ``smoothed_aggregation``
(``src/interpolation/mod.rs:730-836``) by composing the C pieces.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfamg_oracle.so")
_SRC = os.path.join(_HERE, "famg_oracle.c")


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc; a few seconds)."""
    stale = (not os.path.exists(_SO)) or (
        os.path.exists(_SRC) and os.path.getmtime(_SRC) > os.path.getmtime(_SO)
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None
i64p = C.POINTER(C.c_int64)
f64p = C.POINTER(C.c_double)


class _CgInfo(C.Structure):
    _fields_ = [("iters", C.c_int64), ("abs_res", C.c_double), ("rel_res", C.c_double),
                ("status", C.c_int)]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = C.CDLL(_SO)
    vp = C.c_void_p
    sig = {
        "orc_csr_new": (vp, [C.c_int64, C.c_int64, i64p, i64p, f64p]),
        "orc_csr_free": (None, [vp]),
        "orc_csr_nrows": (C.c_int64, [vp]),
        "orc_csr_ncols": (C.c_int64, [vp]),
        "orc_csr_nnz": (C.c_int64, [vp]),
        "orc_csr_row_ptr": (i64p, [vp]),
        "orc_csr_col": (i64p, [vp]),
        "orc_csr_val": (f64p, [vp]),
        "orc_csr_from_triplets": (vp, [C.c_int64, C.c_int64, C.c_int64, i64p, i64p, f64p]),
        "orc_gen_g1": (vp, [C.c_int64]),
        "orc_gen_g1_interp": (vp, [C.c_int64]),
        "orc_gen_g1_restrict": (vp, [C.c_int64]),
        "orc_gen_g7": (vp, [C.c_int64, C.c_int64, C.c_int64]),
        "orc_gen_g27": (vp, [C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_double]),
        "orc_spmm_csr": (None, [vp, f64p, C.c_int64, f64p, C.c_int64, C.c_int64]),
        "orc_parspmm_new": (vp, [vp, C.c_int]),
        "orc_parspmm_free": (None, [vp]),
        "orc_parspmm_apply": (None, [vp, f64p, C.c_int64, f64p, C.c_int64, C.c_int64]),
        "orc_new_l1": (C.c_int, [vp, f64p]),
        "orc_new_l2": (C.c_int, [vp, f64p]),
        "orc_new_jacobi": (C.c_int, [vp, C.c_double, f64p]),
        "orc_spgemm": (vp, [vp, vp]),
        "orc_transpose": (vp, [vp]),
        "orc_smooth_interpolation": (vp, [vp, vp, C.c_double]),
        "orc_thin_q": (None, [C.c_int64, C.c_int64, f64p, C.c_int64]),
        "orc_tentative_p": (vp, [C.c_int64, C.c_int64, C.c_int64, C.c_int64, f64p, C.c_int64,
                                 C.c_int64, i64p, i64p, f64p]),
        "orc_mg_new": (vp, []),
        "orc_mg_add_level": (C.c_int, [vp, vp, C.c_int, f64p, C.c_int64, i64p, i64p, vp, vp, C.c_int]),
        "orc_mg_set_cycle": (None, [vp, C.c_int, C.c_int]),
        "orc_mg_free": (None, [vp]),
        "orc_mg_apply": (None, [vp, f64p, f64p, C.c_int64]),
        "orc_smooth_diag": (None, [vp, f64p, f64p, f64p, C.c_int64, C.c_int]),
        "orc_block_smoother_apply": (C.c_int, [vp, C.c_int64, i64p, i64p, f64p, C.c_int64]),
        "orc_stationary_iteration": (None, [vp, f64p, C.c_int, f64p, C.c_int64]),
        "orc_pcg": (C.c_int, [vp, vp, C.c_int, f64p, vp, f64p, f64p, C.c_double, C.c_double,
                              C.c_int64, C.c_int, C.POINTER(_CgInfo)]),
        "orc_stationary_solver": (C.c_int64, [vp, C.c_int, f64p, vp, f64p, f64p, C.c_int64, C.c_double]),
        "orc_num_threads": (C.c_int, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _pi(a):
    return a.ctypes.data_as(i64p)


def _pf(a):
    return a.ctypes.data_as(f64p) if a is not None else None


def _fcol(x):
    """Column-major n x k float64 array (faer Mat layout)."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    return np.asfortranarray(x)


class Csr:
    """Host CSR with the reference's index width (usize -> int64). Owns a C-side copy."""

    def __init__(self, handle):
        if not handle:
            raise ValueError("oracle returned NULL (the reference would have panicked here)")
        self._h = handle
        L = lib()
        self.nrows = L.orc_csr_nrows(handle)
        self.ncols = L.orc_csr_ncols(handle)
        self.nnz = L.orc_csr_nnz(handle)
        self.row_ptr = np.ctypeslib.as_array(L.orc_csr_row_ptr(handle), shape=(self.nrows + 1,))
        self.col = np.ctypeslib.as_array(L.orc_csr_col(handle), shape=(max(self.nnz, 1),))[: self.nnz]
        self.val = np.ctypeslib.as_array(L.orc_csr_val(handle), shape=(max(self.nnz, 1),))[: self.nnz]

    def __del__(self):
        try:
            if self._h:
                lib().orc_csr_free(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def shape(self):
        return (self.nrows, self.ncols)

    @staticmethod
    def from_arrays(nrows, ncols, row_ptr, col, val) -> "Csr":
        rp, c, v = _i64(row_ptr), _i64(col), _f64(val)
        return Csr(lib().orc_csr_new(nrows, ncols, _pi(rp), _pi(c), _pf(v)))

    @staticmethod
    def from_triplets(nrows, ncols, rows, cols, vals) -> "Csr":
        r, c, v = _i64(rows), _i64(cols), _f64(vals)
        return Csr(lib().orc_csr_from_triplets(nrows, ncols, len(r), _pi(r), _pi(c), _pf(v)))

    @staticmethod
    def from_scipy(m) -> "Csr":
        m = m.tocsr()
        m.sort_indices()
        return Csr.from_arrays(m.shape[0], m.shape[1], m.indptr, m.indices, m.data)

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.val.copy(), self.col.copy(), self.row_ptr.copy()), shape=self.shape)


# ----------------------------------------------------------------------------- generators
def gen_g1(n_elements: int) -> Csr:
    return Csr(lib().orc_gen_g1(n_elements))


def gen_g1_interp(n_coarse: int) -> Csr:
    return Csr(lib().orc_gen_g1_interp(n_coarse))


def gen_g1_restrict(n_coarse: int) -> Csr:
    return Csr(lib().orc_gen_g1_restrict(n_coarse))


def gen_g7(nx: int, ny: Optional[int] = None, nz: Optional[int] = None) -> Csr:
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    return Csr(lib().orc_gen_g7(nx, ny, nz))


def gen_g27(nx: int, ny: Optional[int] = None, nz: Optional[int] = None, ey: float = 1.0,
            ez: float = 1e-2) -> Csr:
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    return Csr(lib().orc_gen_g27(nx, ny, nz, ey, ez))


def geometric_aggregates(dims: Sequence[int], block: Sequence[int] = (2, 2, 2)):
    """Deterministic geometric aggregates: bx x by x bz boxes of a lexicographic
    grid (i = x + nx*(y + ny*z)); a trailing partial box is merged into its predecessor so no
    aggregate is smaller than a full box along that axis.  Returns (agg_ptr, agg_nodes,
    coarse_dims) with nodes ascending inside each aggregate (BTreeSet order)."""
    nx, ny, nz = dims
    bx, by, bz = block
    cx, cy, cz = max(nx // bx, 1), max(ny // by, 1), max(nz // bz, 1)
    x = np.minimum(np.arange(nx) // bx, cx - 1)
    y = np.minimum(np.arange(ny) // by, cy - 1)
    z = np.minimum(np.arange(nz) // bz, cz - 1)
    agg = (x[None, None, :] + cx * (y[None, :, None] + cy * z[:, None, None])).reshape(-1)
    order = np.argsort(agg, kind="stable").astype(np.int64)
    counts = np.bincount(agg, minlength=cx * cy * cz)
    agg_ptr = np.zeros(cx * cy * cz + 1, dtype=np.int64)
    np.cumsum(counts, out=agg_ptr[1:])
    return agg_ptr, order, (cx, cy, cz)


# ----------------------------------------------------------------------------- kernels
def spmm_csr(a: Csr, x) -> np.ndarray:
    x = _fcol(x)
    y = np.zeros((a.nrows, x.shape[1]), order="F")
    lib().orc_spmm_csr(a._h, _pf(x), x.shape[0], _pf(y), a.nrows, x.shape[1])
    return y


class ParSpmmOp:
    """par_spmm.rs ParSpmmOp restated (tiled, OpenMP over block-rows)."""

    def __init__(self, a: Csr, nthreads: int):
        self.a = a
        self._h = lib().orc_parspmm_new(a._h, nthreads)
        if not self._h:
            raise ValueError("not a parallel operator with a single thread")  # par_spmm.rs:35
        self.nthreads = nthreads

    def apply(self, x) -> np.ndarray:
        x = _fcol(x)
        y = np.empty((self.a.nrows, x.shape[1]), order="F")
        lib().orc_parspmm_apply(self._h, _pf(x), x.shape[0], _pf(y), self.a.nrows, x.shape[1])
        return y

    def __del__(self):
        try:
            lib().orc_parspmm_free(self._h)
        except Exception:
            pass


def new_l1(a: Csr) -> np.ndarray:
    d = np.empty(a.nrows)
    lib().orc_new_l1(a._h, _pf(d))
    return d


def new_l2(a: Csr) -> np.ndarray:
    d = np.empty(a.nrows)
    if lib().orc_new_l2(a._h, _pf(d)):
        raise ValueError("missing diagonal entry")
    return d


def new_jacobi(a: Csr, omega: float) -> np.ndarray:
    d = np.empty(a.nrows)
    if lib().orc_new_jacobi(a._h, omega, _pf(d)):
        raise ValueError("missing diagonal entry")
    return d


def spgemm(a: Csr, b: Csr) -> Csr:
    return Csr(lib().orc_spgemm(a._h, b._h))


def transpose(a: Csr) -> Csr:
    return Csr(lib().orc_transpose(a._h))


def smooth_interpolation(a: Csr, p: Csr, omega: float = 0.66) -> Csr:
    return Csr(lib().orc_smooth_interpolation(a._h, p._h, omega))


def thin_q(m) -> np.ndarray:
    m = _fcol(m).copy(order="F")
    lib().orc_thin_q(m.shape[0], m.shape[1], _pf(m), m.shape[0])
    return m


def tentative_p(n_fine, near_null, agg_ptr, agg_nodes, cand: int = 1, block_size: int = 1):
    nn = _fcol(near_null)
    agg_ptr, agg_nodes = _i64(agg_ptr), _i64(agg_nodes)
    n_aggs = len(agg_ptr) - 1
    coarse_nn = np.zeros((n_aggs * cand, nn.shape[1]), order="F")
    h = lib().orc_tentative_p(n_fine, block_size, nn.shape[1], cand, _pf(nn), nn.shape[0], n_aggs,
                              _pi(agg_ptr), _pi(agg_nodes), _pf(coarse_nn))
    return Csr(h), coarse_nn


def smooth_diag(a: Csr, d, x, b, iters: int = 1) -> np.ndarray:
    x = _fcol(x).copy(order="F")
    b = _fcol(b)
    d = _f64(d)
    lib().orc_smooth_diag(a._h, _pf(d), _pf(x), _pf(b), x.shape[1], iters)
    return x


def block_smoother_apply(a: Csr, agg_ptr, agg_nodes, r) -> np.ndarray:
    r = _fcol(r).copy(order="F")
    agg_ptr, agg_nodes = _i64(agg_ptr), _i64(agg_nodes)
    if lib().orc_block_smoother_apply(a._h, len(agg_ptr) - 1, _pi(agg_ptr), _pi(agg_nodes), _pf(r),
                                      r.shape[1]):
        raise ValueError("block factorisation failed")
    return r


def stationary_iteration(a: Csr, d, iters: int, x) -> np.ndarray:
    x = _fcol(x).copy(order="F")
    d = _f64(d)
    lib().orc_stationary_iteration(a._h, _pf(d), iters, _pf(x), x.shape[1])
    return x


def stationary_iteration_transpose(a: Csr, d, iters: int, rhs) -> np.ndarray:
    """StationaryIteration::transpose_apply (smoothers.rs:179-197) with a Diag preconditioner."""
    w1 = _fcol(rhs).copy(order="F")
    d = _f64(d).reshape(-1, 1)
    for _ in range(iters):
        w2 = d * w1
        out = spmm_csr(a, w2)
        w1 = w1 - out
    return w1


def error_propagator(a: Csr, precond_apply, x) -> np.ndarray:
    """ErrorPropogator::apply (adaptivity.rs:191-198): x - M^-1 (A x)."""
    x = _fcol(x)
    return x - precond_apply(spmm_csr(a, x))


def _add_assign_pattern(s: Csr, p: Csr) -> Csr:
    """faer ``add_assign(smoothed.transpose_mut(), p.transpose())`` [faer-recalled]: S += P entrywise;
    pattern(P) must be contained in pattern(S) (panics otherwise)."""
    val = s.val.copy()
    for i in range(p.nrows):
        cols = s.col[s.row_ptr[i]:s.row_ptr[i + 1]]
        for q in range(p.row_ptr[i], p.row_ptr[i + 1]):
            u = int(np.searchsorted(cols, p.col[q]))
            if u == len(cols) or cols[u] != p.col[q]:
                raise ValueError("pattern(P) is not contained in pattern(S)")
            val[s.row_ptr[i] + u] += p.val[q]
    return Csr.from_arrays(s.nrows, s.ncols, s.row_ptr, s.col, val)


def block_diag_inverse(a: Csr, block_size: int, scale: float) -> Csr:
    """The D^-1 of block_jacobi (interpolation/mod.rs:968-1015): per diagonal block,
    ``self_adjoint_eigen(Lower)`` -> U diag(1/s) U^T (eigenvalues must exceed 1e-6), every entry kept as
    a triplet, scaled.  LAPACK's ``eigh`` (lower side) stands in for faer's."""
    n = a.nrows
    rows, cols, vals = [], [], []
    dense_get = a.to_scipy().tocsr()
    for b in range(n // block_size):
        st = b * block_size
        blk = dense_get[st:st + block_size, st:st + block_size].toarray()
        w, u = np.linalg.eigh(blk, UPLO="L")
        if np.any(w <= 1e-6):
            raise ValueError(f"block diagonal is nearly singular with eigval of: {w.min():.3e}")
        inv = (u * (1.0 / w)) @ u.T
        for i in range(block_size):
            for j in range(block_size):
                rows.append(st + i); cols.append(st + j); vals.append(scale * inv[i, j])
    return Csr.from_triplets(n, n, rows, cols, vals)


def block_jacobi(a: Csr, block_size: int, p: Csr) -> Csr:
    """block_jacobi (interpolation/mod.rs:963-1028): smoothed = d_inv * (mat * p); smoothed += p."""
    d_inv = block_diag_inverse(a, block_size, -0.66)
    return _add_assign_pattern(spgemm(d_inv, spgemm(a, p)), p)


def smooth_p(a: Csr, m_inv: Csr, p: Csr) -> Csr:
    """smooth_p (interpolation/mod.rs:1030-1040): ap = mat * p; ap *= -1; smoothed = m_inv * ap; += p."""
    ap = spgemm(a, p)
    ap = Csr.from_arrays(ap.nrows, ap.ncols, ap.row_ptr, ap.col, -ap.val)
    return _add_assign_pattern(spgemm(m_inv, ap), p)


def block_smoother_vector_apply(a: Csr, vdim: int, agg_ptr, agg_nodes, r) -> np.ndarray:
    """BlockSmoother::apply for vdim > 1 (block_smoothers.rs:165-214) with diagonally_compensate_vector
    (:326-400): per aggregate of nodes, in-aggregate couplings kept, every coupling block to an outside
    node lumped into the node's diagonal block as 0.5 U S U^T (-A_IJ = U S V^T, LAPACK svd), block solved
    exactly (lower triangle, as a Cholesky of one side would)."""
    r = _fcol(r).copy(order="F")
    m = a.to_scipy().tocsr()
    agg_ptr, agg_nodes = np.asarray(agg_ptr), np.asarray(agg_nodes)
    nnodes = a.nrows // vdim
    node_agg = np.empty(nnodes, dtype=np.int64)
    for g in range(len(agg_ptr) - 1):
        node_agg[agg_nodes[agg_ptr[g]:agg_ptr[g + 1]]] = g
    for g in range(len(agg_ptr) - 1):
        nodes = agg_nodes[agg_ptr[g]:agg_ptr[g + 1]]
        dofs = (nodes[:, None] * vdim + np.arange(vdim)[None, :]).reshape(-1)
        blk = m[dofs][:, dofs].toarray()
        for li, bi in enumerate(nodes):
            rows = m[bi * vdim:(bi + 1) * vdim]
            outside = sorted({int(j) // vdim for j in rows.indices if node_agg[int(j) // vdim] != g})
            for bj in outside:
                aij = -rows[:, bj * vdim:(bj + 1) * vdim].toarray()
                u, sv, _ = np.linalg.svd(aij)
                blk[li * vdim:(li + 1) * vdim, li * vdim:(li + 1) * vdim] += 0.5 * ((u * sv) @ u.T)
        low = np.tril(blk)
        sym = low + np.tril(blk, -1).T
        r[dofs, :] = np.linalg.solve(sym, r[dofs, :])
    return r


def composite_apply(a: Csr, components, rhs) -> np.ndarray:
    """Composite::implementation (composite.rs:66-83): ``components`` are callables ws -> M^-1 ws."""
    rhs = _fcol(rhs)
    out = np.zeros_like(rhs, order="F")
    ws = rhs.copy(order="F")
    for comp in reversed(components):
        ws = comp(ws)
        out = out + ws
        ws = rhs - spmm_csr(a, out)
    for comp in components[1:]:
        ws = comp(ws)
        out = out + ws
        ws = rhs - spmm_csr(a, out)
    return out


def smooth_vector(a: Csr, precond_apply, x0, iterations: int):
    """smooth_vector (adaptivity.rs:307-390) from a given start block x0 (the reference draws it from
    an unseeded StandardNormal stream, :321-329): x = thinQ(thinQ(x0)); iterations x { x = E x;
    x = thinQ(x) } with the Householder thin Q of ``orc_thin_q``; then per column
    cf = ||E w||_A / ||w||_A (:365-384).  Returns (x, cfs)."""
    x = thin_q(thin_q(_fcol(x0)))
    for _ in range(iterations):
        x = thin_q(error_propagator(a, precond_apply, x))
    cfs = []
    for c in range(x.shape[1]):
        w = x[:, c:c + 1]
        aw = spmm_csr(a, w)
        w_a = np.sqrt(float(w[:, 0] @ aw[:, 0]))
        ev = w - precond_apply(aw)
        aev = spmm_csr(a, ev)
        cfs.append(np.sqrt(float(ev[:, 0] @ aev[:, 0])) / w_a)
    return x, cfs


def create_weights(a: Csr, nn_basis) -> list:
    """create_weights (adaptivity.rs:434-443): 1 / (v^T A v) per column."""
    v = _fcol(nn_basis)
    av = spmm_csr(a, v)
    return [1.0 / float(v[:, c] @ av[:, c]) for c in range(v.shape[1])]


# ----------------------------------------------------------------------------- multigrid
SM_DIAG, SM_LLT, SM_BLOCK = 0, 1, 2
PC_NONE, PC_DIAG, PC_MG = 0, 1, 2


class Multigrid:
    """multigrid.rs Multigrid restated: new / add_level / with_cycle_type / with_smoothing_steps /
    apply."""

    def __init__(self, nthreads: int = 1):
        self._h = lib().orc_mg_new()
        self._keep = []
        self.nthreads = nthreads
        self.levels = 0

    def add_level(self, a: Csr, smoother, r: Optional[Csr] = None, p: Optional[Csr] = None):
        """smoother: ndarray (Diag), the string 'cholesky', or ('block', agg_ptr, agg_nodes)."""
        L = lib()
        self._keep += [a, r, p]
        rh = r._h if r is not None else None
        ph = p._h if p is not None else None
        if isinstance(smoother, str) and smoother == "cholesky":
            rc = L.orc_mg_add_level(self._h, a._h, SM_LLT, None, 0, None, None, rh, ph, self.nthreads)
        elif isinstance(smoother, tuple):
            ap, an = _i64(smoother[1]), _i64(smoother[2])
            rc = L.orc_mg_add_level(self._h, a._h, SM_BLOCK, None, len(ap) - 1, _pi(ap), _pi(an), rh, ph,
                                    self.nthreads)
        else:
            d = _f64(smoother)
            rc = L.orc_mg_add_level(self._h, a._h, SM_DIAG, _pf(d), 0, None, None, rh, ph, self.nthreads)
        if rc < 0:
            raise ValueError("add_level: shape mismatch or factorisation failure")
        self.levels += 1
        self.n = a.nrows if self.levels == 1 else self.n
        return self

    def with_cycle(self, mu: int = 1, nu: int = 1):
        assert mu > 0 and nu > 0
        lib().orc_mg_set_cycle(self._h, mu, nu)
        return self

    def apply(self, rhs) -> np.ndarray:
        rhs = _fcol(rhs)
        out = np.empty_like(rhs, order="F")
        lib().orc_mg_apply(self._h, _pf(out), _pf(rhs), rhs.shape[1])
        return out

    def __del__(self):
        try:
            lib().orc_mg_free(self._h)
        except Exception:
            pass


@dataclass
class CgInfo:
    iters: int
    abs_residual: float
    rel_residual: float
    status: int  # 0 converged, 1 NoConvergence, 2 not positive definite


def pcg(a: Csr, b, precond=None, x0=None, rel_tol=1e-8, abs_tol=0.0, max_iters=1000,
        par: Optional[ParSpmmOp] = None) -> Tuple[np.ndarray, CgInfo]:
    """precond: None | ndarray (Diag) | Multigrid."""
    b = _f64(b).reshape(-1)
    x = np.zeros_like(b) if x0 is None else _f64(x0).reshape(-1).copy()
    info = _CgInfo()
    kind, d, mg = PC_NONE, None, None
    if isinstance(precond, Multigrid):
        kind, mg = PC_MG, precond._h
    elif precond is not None:
        kind, d = PC_DIAG, _f64(precond)
    lib().orc_pcg(a._h, par._h if par is not None else None, kind, _pf(d), mg, _pf(b), _pf(x),
                  rel_tol, abs_tol, max_iters, 1 if x0 is None else 0, C.byref(info))
    return x, CgInfo(info.iters, info.abs_res, info.rel_res, info.status)


def stationary_solver(a: Csr, b, precond, max_iter: int, rel_tol: float) -> Tuple[np.ndarray, int]:
    b = _f64(b).reshape(-1)
    x = np.zeros_like(b)
    kind, d, mg = PC_NONE, None, None
    if isinstance(precond, Multigrid):
        kind, mg = PC_MG, precond._h
    elif precond is not None:
        kind, d = PC_DIAG, _f64(precond)
    it = lib().orc_stationary_solver(a._h, kind, _pf(d), mg, _pf(b), _pf(x), max_iter, rel_tol)
    return x, int(it)


# ----------------------------------------------------------------------------- hierarchy
@dataclass
class GalerkinCoarse:
    """interpolation/mod.rs:34-40."""
    interpolation: Csr
    restriction: Csr
    coarse_mat: Csr
    coarse_nn: np.ndarray
    partition: Tuple[np.ndarray, np.ndarray]


def smoothed_aggregation(a: Csr, agg_ptr, agg_nodes, near_null, cand: int = 1,
                         smoothing_steps: int = 1, omega: float = 0.66, block_size: int = 1) -> GalerkinCoarse:
    """interpolation/mod.rs:730-836: tentative P by per-aggregate thin SVD, `smoothing_steps` x
    smooth_interpolation(A, P, 0.66) (block_size 1) or block_jacobi (:812-818), R = P^T, A_c = R (A P)."""
    p, coarse_nn = tentative_p(a.nrows, near_null, agg_ptr, agg_nodes, cand, block_size)
    for _ in range(smoothing_steps):
        p = smooth_interpolation(a, p, omega) if block_size == 1 else block_jacobi(a, block_size, p)
    r = transpose(p)
    ac = spgemm(r, spgemm(a, p))
    return GalerkinCoarse(p, r, ac, coarse_nn, (np.asarray(agg_ptr), np.asarray(agg_nodes)))


@dataclass
class Hierarchy:
    """hierarchy.rs:61-70 (operators / restrictions / interpolations / partitions / near_nulls)."""
    operators: List[Csr] = field(default_factory=list)
    restrictions: List[Csr] = field(default_factory=list)
    interpolations: List[Csr] = field(default_factory=list)
    partitions: list = field(default_factory=list)
    near_nulls: List[np.ndarray] = field(default_factory=list)
    grid_dims: list = field(default_factory=list)

    @property
    def levels(self):
        return len(self.operators)

    def op_complexity(self):  # hierarchy.rs:352-360
        return sum(o.nnz for o in self.operators) / self.operators[0].nnz

    def grid_complexity(self):  # hierarchy.rs:346-350
        return sum(o.nrows for o in self.operators) / self.operators[0].nrows


def build_hierarchy(a: Csr, near_null, dims: Sequence[int], coarsest_dim: int = 1000,
                    max_levels: Optional[int] = None, cand: int = 1, smoothing_steps: int = 1,
                    block: Sequence[int] = (2, 2, 2), partitioner=None, block_size: int = 1) -> Hierarchy:
    """Hierarchy::coarsen (hierarchy.rs:190-248) with the partitioner replaced by deterministic
    geometric aggregates (the reference partitioner is non-deterministic) -- or by
    ``partitioner(level, fine, near_null) -> node_to_agg`` (e.g. the restated algebraic partitioner
    of ``oracle/partitioner.py`` with its documented tie-breaking).  Per level:
    GalerkinCoarse; coarse near-null smoothed by a 3-step L1 StationaryIteration (:217-226) and
    re-orthonormalised by thin QR (:228)."""
    h = Hierarchy([a], [], [], [], [_fcol(near_null)], [tuple(dims)])
    level, coarse_dim = 1, np.iinfo(np.int64).max
    max_levels = max_levels if max_levels is not None else np.iinfo(np.int64).max
    while coarse_dim > coarsest_dim and level < max_levels:
        fine = h.operators[-1]
        dims_f = h.grid_dims[-1]
        if partitioner is None:
            agg_ptr, agg_nodes, dims_c = geometric_aggregates(dims_f, block)
        else:
            n2a = np.asarray(partitioner(level - 1, fine, h.near_nulls[-1]), dtype=np.int64)
            agg_nodes = np.argsort(n2a, kind="stable").astype(np.int64)
            agg_ptr = np.concatenate([[0], np.cumsum(np.bincount(n2a))]).astype(np.int64)
            dims_c = None
        # SparseMatOp block size: the caller's on the finest level, candidate_dimension below (hierarchy.rs:210-215)
        bs = block_size if level == 1 else cand
        g = smoothed_aggregation(fine, agg_ptr, agg_nodes, h.near_nulls[-1], cand, smoothing_steps, block_size=bs)
        coarse_dim = g.coarse_mat.nrows
        nn = stationary_iteration(g.coarse_mat, new_l1(g.coarse_mat), 3, g.coarse_nn)
        nn = thin_q(nn)
        h.operators.append(g.coarse_mat)
        h.partitions.append(g.partition)
        h.restrictions.append(g.restriction)
        h.interpolations.append(g.interpolation)
        h.near_nulls.append(nn)
        h.grid_dims.append(dims_c)
        level += 1
    return h


def multigrid_from_hierarchy(h: Hierarchy, smoother: str = "l1", omega: float = 0.66, mu: int = 1,
                             nu: int = 1, nthreads: int = 1) -> Multigrid:
    """Manual Multigrid::new/add_level assembly (the only way the reference wires a diagonal
    smoother, simple_geometric.rs:204-224) with an exact coarsest solve (multigrid.rs:105-109)."""
    mg = Multigrid(nthreads)
    for lvl, a in enumerate(h.operators):
        last = lvl == h.levels - 1
        if last and h.levels > 1:
            sm = "cholesky"
        elif smoother == "l1":
            sm = new_l1(a)
        elif smoother == "l2":
            sm = new_l2(a)
        elif smoother == "jacobi":
            sm = new_jacobi(a, omega)
        else:
            raise ValueError(smoother)
        if lvl == 0:
            mg.add_level(a, sm)
        else:
            mg.add_level(a, sm, h.restrictions[lvl - 1], h.interpolations[lvl - 1])
    return mg.with_cycle(mu, nu)


def mats_are_equal(left: Csr, right: Csr, tol: float = 1e-12) -> bool:
    """utils::mats_are_equal (utils.rs:32-58) with max(|l|,|r|) guarding negative values
    (SURVEY 4): same shape, nnz, identical (row, col) sequence, |d| <= tol abs and rel."""
    if left.shape != right.shape or left.nnz != right.nnz:
        return False
    if not (np.array_equal(left.row_ptr, right.row_ptr) and np.array_equal(left.col, right.col)):
        return False
    absd = np.abs(left.val - right.val)
    den = np.maximum(np.abs(left.val), np.abs(right.val))
    rel = np.where(den > 0, absd / np.where(den > 0, den, 1.0), 0.0)
    return bool(np.all(absd <= tol) and np.all(rel <= tol))


def num_threads() -> int:
    return int(lib().orc_num_threads())
