"""Shared helpers of the parity tests: oracle <-> product conversions and tolerances."""
import numpy as np

import oracle as O


def to_dev(ctx, a: O.Csr):
    import faer_amg_b200 as F
    return F.SparseRowMat.from_csr(ctx, a.nrows, a.ncols, a.row_ptr, a.col, a.val)


def to_oracle(m) -> O.Csr:
    rp, ci, v = m.to_host()
    return O.Csr.from_arrays(m.nrows, m.ncols, rp.astype(np.int64), ci.astype(np.int64), v)


def same_pattern(dev, orc: O.Csr) -> bool:
    rp, ci, _ = dev.to_host()
    return (dev.shape == orc.shape and np.array_equal(rp.astype(np.int64), orc.row_ptr)
            and np.array_equal(ci.astype(np.int64), orc.col))


def spmv_bound(a: O.Csr, x) -> np.ndarray:
    """(|A| |x|)_i: the magnitude a row's rounding error is relative to (summation order)."""
    absa = O.Csr.from_arrays(a.nrows, a.ncols, a.row_ptr, a.col, np.abs(a.val))
    return O.spmm_csr(absa, np.abs(x))


def assert_rel(got, want, bound, tol=1e-12):
    err = np.abs(got - want)
    lim = tol * np.maximum(bound, np.finfo(float).tiny)
    assert np.all(err <= lim), f"max err/bound = {np.max(err / np.maximum(bound, 1e-300)):.3e} > {tol}"


def random_csr(rng, nrows, ncols, row_lens) -> O.Csr:
    """Random CSR with prescribed row lengths (sorted unique columns), values N(0,1)."""
    rp = np.zeros(nrows + 1, dtype=np.int64)
    cols, vals = [], []
    for i in range(nrows):
        k = int(min(row_lens[i], ncols))
        c = np.sort(rng.choice(ncols, size=k, replace=False)) if k else np.zeros(0, dtype=np.int64)
        cols.append(c)
        vals.append(rng.standard_normal(k))
        rp[i + 1] = rp[i] + k
    col = np.concatenate(cols) if cols else np.zeros(0, dtype=np.int64)
    val = np.concatenate(vals) if vals else np.zeros(0)
    return O.Csr.from_arrays(nrows, ncols, rp, col, val)
