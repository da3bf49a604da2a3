"""Synthetic benchmark operators generated on the device (the reference ships no 3-D
generator, F8)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from ._ffi import call, vp
from .core import Context, SparseRowMat


def poisson7(ctx: Context, nx: int, ny: Optional[int] = None, nz: Optional[int] = None) -> SparseRowMat:
    """G7: 7-point Laplacian, lexicographic, diagonal 6 / off-diagonals -1, Dirichlet eliminated."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    h = vp()
    call("famg_gallery_g7", ctx._h, nx, ny, nz, C.byref(h))
    return SparseRowMat(ctx, h)


def diffusion27(ctx: Context, nx: int, ny: Optional[int] = None, nz: Optional[int] = None, eps_y: float = 1.0,
                eps_z: float = 1e-2) -> SparseRowMat:
    """G27: 27-point trilinear-FE anisotropic diffusion Kx(x)My(x)Mz + ey Mx(x)Ky(x)Mz + ez Mx(x)My(x)Kz."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    h = vp()
    call("famg_gallery_g27", ctx._h, nx, ny, nz, float(eps_y), float(eps_z), C.byref(h))
    return SparseRowMat(ctx, h)


def poisson7_slab(ctx: Context, nx: int, ny: int, nz: int, z0: int, z1: int) -> SparseRowMat:
    """Rows of the planes [z0, z1) of G7 with global column ids (one rank's slab of the partitioned operator)."""
    h = vp()
    call("famg_gallery_g7_slab", ctx._h, nx, ny, nz, z0, z1, C.byref(h))
    return SparseRowMat(ctx, h)


def diffusion27_slab(ctx: Context, nx: int, ny: int, nz: int, z0: int, z1: int, eps_y: float = 1.0, eps_z: float = 1e-2) -> SparseRowMat:
    """Rows of the planes [z0, z1) of G27 with global column ids."""
    h = vp()
    call("famg_gallery_g27_slab", ctx._h, nx, ny, nz, float(eps_y), float(eps_z), z0, z1, C.byref(h))
    return SparseRowMat(ctx, h)


def poisson1d(ctx: Context, n_elements: int) -> SparseRowMat:
    """make_finite_difference (examples/simple_geometric.rs:96-113)."""
    h = 1.0 / n_elements
    n = n_elements - 1
    rows, cols, vals = [], [], []
    for i in range(n):
        rows.append(i); cols.append(i); vals.append(2.0 / (h * h))
        if i > 0:
            rows.append(i); cols.append(i - 1); vals.append(-1.0 / (h * h))
        if i + 1 < n:
            rows.append(i); cols.append(i + 1); vals.append(-1.0 / (h * h))
    return SparseRowMat.try_new_from_triplets(ctx, n, n, rows, cols, vals)


def interpolation1d(ctx: Context, n_coarse: int) -> SparseRowMat:
    """make_interpolation (simple_geometric.rs:62-77)."""
    rows, cols, vals = [], [], []
    for c in range(n_coarse):
        for k, w in enumerate((1.0, 2.0, 1.0)):
            rows.append(2 * c + k); cols.append(c); vals.append(w * 0.5)
    return SparseRowMat.try_new_from_triplets(ctx, 2 * n_coarse + 1, n_coarse, rows, cols, vals)


def restriction1d(ctx: Context, n_coarse: int) -> SparseRowMat:
    """make_restriction (simple_geometric.rs:82-93)."""
    rows, cols, vals = [], [], []
    for r in range(n_coarse):
        for k, w in enumerate((1.0, 2.0, 1.0)):
            rows.append(r); cols.append(2 * r + k); vals.append(w * 0.25)
    return SparseRowMat.try_new_from_triplets(ctx, n_coarse, 2 * n_coarse + 1, rows, cols, vals)
