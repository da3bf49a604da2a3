/*
 * famg_oracle.c -- CPU restatement of the faer-amg hot path.   TEST INFRASTRUCTURE ONLY.
 *
 * This file is the *checker*, not the product: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  Nothing under faer_amg_b200/
 * links, imports or calls it.
 *
 * PARITY UNPINNED: the reference (aujxn/faer-amg, Rust) ships no tests, no golden vectors and
 * cannot be compiled here (no rustc; its arithmetic lives in the out-of-tree path dependency
 * faer 0.23.2, Cargo.toml:9-10, Cargo.lock:590-591).  Every function below restates the
 * reference source it cites, plus faer semantics recalled from faer's public source
 * ("[faer-recalled]").  It is cross-checked against scipy and analytic identities in
 * tests/test_oracle.py, but never against the reference binary.
 *
 * Index type is int64 (the reference uses usize); values are f64; dense multivectors are
 * column-major with unit row stride (faer Mat<f64>).
 *
 * Build:  make -C oracle      (gcc -O3 -fopenmp -ffp-contract=off -shared)
 * -ffp-contract=off: the reference's scalar kernels are `dst += a*b` in Rust, which never fuses
 * into an FMA without explicit mul_add; the restatement keeps mul and add separately rounded.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t i64;

typedef struct {
    i64 nrows, ncols;
    i64 *row_ptr; /* nrows+1 */
    i64 *col;     /* nnz, sorted & unique per row */
    double *val;  /* nnz */
} orc_csr;

#define ORC_API __attribute__((visibility("default")))

static void *xmalloc(size_t n) {
    void *p = malloc(n ? n : 1);
    if (!p) { fprintf(stderr, "oracle: out of memory (%zu bytes)\n", n); abort(); }
    return p;
}
static void *xcalloc(size_t n, size_t s) {
    void *p = calloc(n ? n : 1, s ? s : 1);
    if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
    return p;
}

/* ------------------------------------------------------------------ CSR container ---------- */

static orc_csr *csr_alloc(i64 nrows, i64 ncols, i64 nnz) {
    orc_csr *m = (orc_csr *)xmalloc(sizeof(orc_csr));
    m->nrows = nrows; m->ncols = ncols;
    m->row_ptr = (i64 *)xcalloc((size_t)nrows + 1, sizeof(i64));
    m->col = (i64 *)xmalloc(sizeof(i64) * (size_t)nnz);
    m->val = (double *)xmalloc(sizeof(double) * (size_t)nnz);
    return m;
}

ORC_API orc_csr *orc_csr_new(i64 nrows, i64 ncols, const i64 *row_ptr, const i64 *col,
                             const double *val) {
    i64 nnz = row_ptr[nrows];
    orc_csr *m = csr_alloc(nrows, ncols, nnz);
    memcpy(m->row_ptr, row_ptr, sizeof(i64) * (size_t)(nrows + 1));
    memcpy(m->col, col, sizeof(i64) * (size_t)nnz);
    memcpy(m->val, val, sizeof(double) * (size_t)nnz);
    return m;
}
ORC_API void orc_csr_free(orc_csr *m) {
    if (!m) return;
    free(m->row_ptr); free(m->col); free(m->val); free(m);
}
ORC_API i64 orc_csr_nrows(const orc_csr *m) { return m->nrows; }
ORC_API i64 orc_csr_ncols(const orc_csr *m) { return m->ncols; }
ORC_API i64 orc_csr_nnz(const orc_csr *m) { return m->row_ptr[m->nrows]; }
ORC_API const i64 *orc_csr_row_ptr(const orc_csr *m) { return m->row_ptr; }
ORC_API const i64 *orc_csr_col(const orc_csr *m) { return m->col; }
ORC_API const double *orc_csr_val(const orc_csr *m) { return m->val; }

/* SparseRowMat::try_new_from_triplets [faer-recalled] (called at interpolation/mod.rs:807,
 * simple_geometric.rs:76,92,112): CSR with sorted unique column indices per row; duplicate
 * (row,col) entries are summed in input order; explicit zeros are kept.
 * Counting sort by row (stable) then stable insertion sort by column inside each row. */
ORC_API orc_csr *orc_csr_from_triplets(i64 nrows, i64 ncols, i64 nt, const i64 *tr, const i64 *tc,
                                       const double *tv) {
    i64 *cnt = (i64 *)xcalloc((size_t)nrows + 1, sizeof(i64));
    for (i64 t = 0; t < nt; ++t) {
        if (tr[t] < 0 || tr[t] >= nrows || tc[t] < 0 || tc[t] >= ncols) { free(cnt); return NULL; }
        cnt[tr[t] + 1]++;
    }
    for (i64 i = 0; i < nrows; ++i) cnt[i + 1] += cnt[i];
    i64 *pc = (i64 *)xmalloc(sizeof(i64) * (size_t)nt);
    double *pv = (double *)xmalloc(sizeof(double) * (size_t)nt);
    i64 *fill = (i64 *)xmalloc(sizeof(i64) * (size_t)(nrows + 1));
    memcpy(fill, cnt, sizeof(i64) * (size_t)(nrows + 1));
    for (i64 t = 0; t < nt; ++t) { i64 q = fill[tr[t]]++; pc[q] = tc[t]; pv[q] = tv[t]; }
    /* stable sort each row by column (rows are short), then merge duplicates */
    i64 *rp = (i64 *)xcalloc((size_t)nrows + 1, sizeof(i64));
    i64 out = 0;
    for (i64 i = 0; i < nrows; ++i) {
        i64 b = cnt[i], e = cnt[i + 1];
        for (i64 a = b + 1; a < e; ++a) {
            i64 c = pc[a]; double v = pv[a]; i64 q = a - 1;
            while (q >= b && pc[q] > c) { pc[q + 1] = pc[q]; pv[q + 1] = pv[q]; --q; }
            pc[q + 1] = c; pv[q + 1] = v;
        }
        i64 rs = out;
        for (i64 a = b; a < e; ++a) {
            if (out > rs && pc[out - 1] == pc[a]) pv[out - 1] += pv[a];
            else { pc[out] = pc[a]; pv[out] = pv[a]; ++out; }
        }
        rp[i + 1] = out;
    }
    orc_csr *m = csr_alloc(nrows, ncols, out);
    memcpy(m->row_ptr, rp, sizeof(i64) * (size_t)(nrows + 1));
    memcpy(m->col, pc, sizeof(i64) * (size_t)out);
    memcpy(m->val, pv, sizeof(double) * (size_t)out);
    free(cnt); free(pc); free(pv); free(fill); free(rp);
    return m;
}

/* ------------------------------------------------------------------ synthetic inputs ------- */

/* make_finite_difference, examples/simple_geometric.rs:96-113: tridiag(-1,2,-1)/h^2,
 * h = 1/n_elements, n_dofs = n_elements-1. */
ORC_API orc_csr *orc_gen_g1(i64 n_elements) {
    double h = 1.0 / (double)n_elements;
    double dv = 2.0 / (h * h), ov = -1.0 / (h * h);
    i64 n = n_elements - 1;
    i64 nnz = 3 * n - 2;
    orc_csr *m = csr_alloc(n, n, nnz > 0 ? nnz : 0);
    i64 q = 0;
    for (i64 i = 0; i < n; ++i) {
        if (i > 0) { m->col[q] = i - 1; m->val[q++] = ov; }
        m->col[q] = i; m->val[q++] = dv;
        if (i + 1 < n) { m->col[q] = i + 1; m->val[q++] = ov; }
        m->row_ptr[i + 1] = q;
    }
    return m;
}
/* make_interpolation, simple_geometric.rs:62-77: n_fine = 2 n_c + 1; column c has
 * 0.5, 1.0, 0.5 at rows 2c, 2c+1, 2c+2. */
ORC_API orc_csr *orc_gen_g1_interp(i64 nc) {
    i64 nf = 2 * nc + 1, nt = 3 * nc;
    i64 *r = (i64 *)xmalloc(sizeof(i64) * (size_t)nt), *c = (i64 *)xmalloc(sizeof(i64) * (size_t)nt);
    double *v = (double *)xmalloc(sizeof(double) * (size_t)nt);
    double scale = 1.0 / 2.0;
    for (i64 j = 0; j < nc; ++j) {
        r[3 * j] = 2 * j; c[3 * j] = j; v[3 * j] = 1.0 * scale;
        r[3 * j + 1] = 2 * j + 1; c[3 * j + 1] = j; v[3 * j + 1] = 2.0 * scale;
        r[3 * j + 2] = 2 * j + 2; c[3 * j + 2] = j; v[3 * j + 2] = 1.0 * scale;
    }
    orc_csr *m = orc_csr_from_triplets(nf, nc, nt, r, c, v);
    free(r); free(c); free(v);
    return m;
}
/* make_restriction, simple_geometric.rs:82-93: full weighting, row r has 0.25,0.5,0.25 at
 * columns 2r, 2r+1, 2r+2. */
ORC_API orc_csr *orc_gen_g1_restrict(i64 nc) {
    i64 nf = 2 * nc + 1, nt = 3 * nc;
    i64 *r = (i64 *)xmalloc(sizeof(i64) * (size_t)nt), *c = (i64 *)xmalloc(sizeof(i64) * (size_t)nt);
    double *v = (double *)xmalloc(sizeof(double) * (size_t)nt);
    double scale = 1.0 / 4.0;
    for (i64 j = 0; j < nc; ++j) {
        r[3 * j] = j; c[3 * j] = 2 * j; v[3 * j] = 1.0 * scale;
        r[3 * j + 1] = j; c[3 * j + 1] = 2 * j + 1; v[3 * j + 1] = 2.0 * scale;
        r[3 * j + 2] = j; c[3 * j + 2] = 2 * j + 2; v[3 * j + 2] = 1.0 * scale;
    }
    orc_csr *m = orc_csr_from_triplets(nc, nf, nt, r, c, v);
    free(r); free(c); free(v);
    return m;
}

/* G7(nx,ny,nz): SURVEY.md 8(d).  7-point Laplacian, lexicographic i = x + nx*(y + ny*z),
 * diagonal 6, off-diagonals -1 to existing neighbours (Dirichlet eliminated). */
ORC_API orc_csr *orc_gen_g7(i64 nx, i64 ny, i64 nz) {
    i64 n = nx * ny * nz;
    i64 nnz = 7 * n - 2 * (nx * ny + ny * nz + nx * nz);
    orc_csr *m = csr_alloc(n, n, nnz);
    i64 *rp = m->row_ptr;
#pragma omp parallel for schedule(static)
    for (i64 i = 0; i < n; ++i) {
        i64 x = i % nx, y = (i / nx) % ny, z = i / (nx * ny);
        rp[i + 1] = 1 + (x > 0) + (x + 1 < nx) + (y > 0) + (y + 1 < ny) + (z > 0) + (z + 1 < nz);
    }
    for (i64 i = 0; i < n; ++i) rp[i + 1] += rp[i];
#pragma omp parallel for schedule(static)
    for (i64 i = 0; i < n; ++i) {
        i64 x = i % nx, y = (i / nx) % ny, z = i / (nx * ny);
        i64 q = rp[i];
        if (z > 0) { m->col[q] = i - nx * ny; m->val[q++] = -1.0; }
        if (y > 0) { m->col[q] = i - nx; m->val[q++] = -1.0; }
        if (x > 0) { m->col[q] = i - 1; m->val[q++] = -1.0; }
        m->col[q] = i; m->val[q++] = 6.0;
        if (x + 1 < nx) { m->col[q] = i + 1; m->val[q++] = -1.0; }
        if (y + 1 < ny) { m->col[q] = i + nx; m->val[q++] = -1.0; }
        if (z + 1 < nz) { m->col[q] = i + nx * ny; m->val[q++] = -1.0; }
    }
    return m;
}

/* G27(N,ey,ez): SURVEY.md 8(d).  Kx(x)My(x)Mz + ey Mx(x)Ky(x)Mz + ez Mx(x)My(x)Kz with
 * K = tridiag(-1,2,-1), M = tridiag(1,4,1)/6.  Entry for offset (dx,dy,dz):
 *   k(dx) m(dy) m(dz) + ey m(dx) k(dy) m(dz) + ez m(dx) m(dy) k(dz)
 * evaluated in exactly this left-to-right order. */
static inline double g27_k(i64 d) { return d == 0 ? 2.0 : -1.0; }
static inline double g27_m(i64 d) { return d == 0 ? 4.0 / 6.0 : 1.0 / 6.0; }
ORC_API orc_csr *orc_gen_g27(i64 nx, i64 ny, i64 nz, double ey, double ez) {
    i64 n = nx * ny * nz;
    i64 *rp = (i64 *)xcalloc((size_t)n + 1, sizeof(i64));
#pragma omp parallel for schedule(static)
    for (i64 i = 0; i < n; ++i) {
        i64 x = i % nx, y = (i / nx) % ny, z = i / (nx * ny);
        i64 cx = 1 + (x > 0) + (x + 1 < nx), cy = 1 + (y > 0) + (y + 1 < ny),
            cz = 1 + (z > 0) + (z + 1 < nz);
        rp[i + 1] = cx * cy * cz;
    }
    for (i64 i = 0; i < n; ++i) rp[i + 1] += rp[i];
    orc_csr *m = csr_alloc(n, n, rp[n]);
    memcpy(m->row_ptr, rp, sizeof(i64) * (size_t)(n + 1));
    free(rp);
#pragma omp parallel for schedule(static)
    for (i64 i = 0; i < n; ++i) {
        i64 x = i % nx, y = (i / nx) % ny, z = i / (nx * ny);
        i64 q = m->row_ptr[i];
        for (i64 dz = -1; dz <= 1; ++dz) {
            if (z + dz < 0 || z + dz >= nz) continue;
            for (i64 dy = -1; dy <= 1; ++dy) {
                if (y + dy < 0 || y + dy >= ny) continue;
                for (i64 dx = -1; dx <= 1; ++dx) {
                    if (x + dx < 0 || x + dx >= nx) continue;
                    double t0 = g27_k(dx) * g27_m(dy) * g27_m(dz);
                    double t1 = ey * g27_m(dx) * g27_k(dy) * g27_m(dz);
                    double t2 = ez * g27_m(dx) * g27_m(dy) * g27_k(dz);
                    m->col[q] = i + dx + nx * (dy + ny * dz);
                    m->val[q++] = t0 + t1 + t2;
                }
            }
        }
    }
    return m;
}

/* ------------------------------------------------------------------ SpMM ------------------- */

/* faer `impl LinOp for SparseRowMat` [faer-recalled] -- the serial fallback handed out by
 * SparseMatOp::dyn_op (core.rs:88-92) and used for R and P (multigrid.rs:157-159).
 * out is overwritten; per output entry the adds run in ascending column order. */
ORC_API void orc_spmm_csr(const orc_csr *a, const double *x, i64 ldx, double *y, i64 ldy, i64 k) {
    for (i64 c = 0; c < k; ++c) {
        const double *xc = x + c * ldx;
        double *yc = y + c * ldy;
        for (i64 i = 0; i < a->nrows; ++i) {
            double s = 0.0;
            for (i64 q = a->row_ptr[i]; q < a->row_ptr[i + 1]; ++q) s += a->val[q] * xc[a->col[q]];
            yc[i] = s;
        }
    }
}

/* ParSpmmOp (par_spmm.rs:15-133).  PAR_BLOCK_SIZE = 8192 (:15).  new(): one task per block-row;
 * entries bucketed by block-column (BTreeMap => ascending block_j) and turned into CSC tiles
 * (:40-84).  F10a: the reference bounds block-rows with mat.ncols() (:46); the evident intent
 * (nrows) is restated here -- identical for the square operators SparseMatOp admits (core.rs:57).
 * apply(): per block-row zero the output slab, then for each tile in ascending block_j call
 * sparse_dense_matmul(Accum::Add, CSC tile, rhs slab, 1.0, Par::Seq) (:115-132), which
 * [faer-recalled] walks rhs columns, then tile columns ascending, scattering
 * dst[i] += tile[i,k]*rhs[k].  Block-rows run in parallel (rayon :100-110 -> OpenMP here). */
#define PAR_BLOCK_SIZE 8192
typedef struct {
    i64 ncols_local;
    i64 block_j;
    i64 *col_ptr;  /* ncols_local+1 */
    int32_t *row;  /* local row */
    double *val;
} orc_tile;
typedef struct {
    i64 ntiles;
    orc_tile *tiles;
} orc_blockrow;
typedef struct {
    i64 nrows, ncols, nblock_rows;
    orc_blockrow *rows;
    int nthreads;
} orc_parspmm;

ORC_API orc_parspmm *orc_parspmm_new(const orc_csr *a, int nthreads) {
    if (nthreads <= 1) return NULL; /* par_spmm.rs:33-36: panics for a single thread */
    orc_parspmm *op = (orc_parspmm *)xmalloc(sizeof(orc_parspmm));
    op->nrows = a->nrows; op->ncols = a->ncols; op->nthreads = nthreads;
    op->nblock_rows = (a->nrows + PAR_BLOCK_SIZE - 1) / PAR_BLOCK_SIZE;
    op->rows = (orc_blockrow *)xcalloc((size_t)op->nblock_rows, sizeof(orc_blockrow));
    i64 nbc = (a->ncols + PAR_BLOCK_SIZE - 1) / PAR_BLOCK_SIZE;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (i64 bi = 0; bi < op->nblock_rows; ++bi) {
        i64 r0 = bi * PAR_BLOCK_SIZE;
        i64 r1 = r0 + PAR_BLOCK_SIZE < a->nrows ? r0 + PAR_BLOCK_SIZE : a->nrows;
        i64 *bcnt = (i64 *)xcalloc((size_t)nbc, sizeof(i64));
        for (i64 i = r0; i < r1; ++i)
            for (i64 q = a->row_ptr[i]; q < a->row_ptr[i + 1]; ++q) bcnt[a->col[q] / PAR_BLOCK_SIZE]++;
        i64 nt = 0;
        for (i64 bj = 0; bj < nbc; ++bj) nt += bcnt[bj] > 0;
        orc_blockrow *br = &op->rows[bi];
        br->ntiles = nt;
        br->tiles = (orc_tile *)xcalloc((size_t)nt, sizeof(orc_tile));
        i64 *tile_of = (i64 *)xmalloc(sizeof(i64) * (size_t)nbc);
        i64 t = 0;
        for (i64 bj = 0; bj < nbc; ++bj) {
            tile_of[bj] = -1;
            if (!bcnt[bj]) continue;
            orc_tile *tl = &br->tiles[t];
            i64 c0 = bj * PAR_BLOCK_SIZE;
            i64 c1 = (bj + 1) * PAR_BLOCK_SIZE < a->ncols ? (bj + 1) * PAR_BLOCK_SIZE : a->ncols;
            tl->block_j = bj; tl->ncols_local = c1 - c0;
            tl->col_ptr = (i64 *)xcalloc((size_t)tl->ncols_local + 1, sizeof(i64));
            tl->row = (int32_t *)xmalloc(sizeof(int32_t) * (size_t)bcnt[bj]);
            tl->val = (double *)xmalloc(sizeof(double) * (size_t)bcnt[bj]);
            tile_of[bj] = t++;
        }
        /* CSC by counting sort on local column; row order inside a column stays ascending
         * because rows are visited ascending (== try_new_from_triplets on the pushed triplets) */
        for (i64 i = r0; i < r1; ++i)
            for (i64 q = a->row_ptr[i]; q < a->row_ptr[i + 1]; ++q) {
                i64 j = a->col[q];
                br->tiles[tile_of[j / PAR_BLOCK_SIZE]].col_ptr[j % PAR_BLOCK_SIZE + 1]++;
            }
        for (i64 u = 0; u < nt; ++u) {
            orc_tile *tl = &br->tiles[u];
            for (i64 c = 0; c < tl->ncols_local; ++c) tl->col_ptr[c + 1] += tl->col_ptr[c];
        }
        i64 **fill = (i64 **)xmalloc(sizeof(i64 *) * (size_t)nt);
        for (i64 u = 0; u < nt; ++u) {
            fill[u] = (i64 *)xmalloc(sizeof(i64) * (size_t)(br->tiles[u].ncols_local + 1));
            memcpy(fill[u], br->tiles[u].col_ptr, sizeof(i64) * (size_t)(br->tiles[u].ncols_local + 1));
        }
        for (i64 i = r0; i < r1; ++i)
            for (i64 q = a->row_ptr[i]; q < a->row_ptr[i + 1]; ++q) {
                i64 j = a->col[q], u = tile_of[j / PAR_BLOCK_SIZE];
                i64 pos = fill[u][j % PAR_BLOCK_SIZE]++;
                br->tiles[u].row[pos] = (int32_t)(i - r0);
                br->tiles[u].val[pos] = a->val[q];
            }
        for (i64 u = 0; u < nt; ++u) free(fill[u]);
        free(fill); free(tile_of); free(bcnt);
    }
    return op;
}
ORC_API void orc_parspmm_free(orc_parspmm *op) {
    if (!op) return;
    for (i64 bi = 0; bi < op->nblock_rows; ++bi) {
        for (i64 u = 0; u < op->rows[bi].ntiles; ++u) {
            free(op->rows[bi].tiles[u].col_ptr); free(op->rows[bi].tiles[u].row);
            free(op->rows[bi].tiles[u].val);
        }
        free(op->rows[bi].tiles);
    }
    free(op->rows); free(op);
}
ORC_API void orc_parspmm_apply(const orc_parspmm *op, const double *x, i64 ldx, double *y, i64 ldy,
                               i64 k) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(op->nthreads)
    for (i64 bi = 0; bi < op->nblock_rows; ++bi) {
        i64 r0 = bi * PAR_BLOCK_SIZE;
        i64 r1 = r0 + PAR_BLOCK_SIZE < op->nrows ? r0 + PAR_BLOCK_SIZE : op->nrows;
        const orc_blockrow *br = &op->rows[bi];
        for (i64 c = 0; c < k; ++c) memset(y + c * ldy + r0, 0, sizeof(double) * (size_t)(r1 - r0));
        for (i64 u = 0; u < br->ntiles; ++u) {
            const orc_tile *tl = &br->tiles[u];
            i64 c0 = tl->block_j * PAR_BLOCK_SIZE;
            for (i64 c = 0; c < k; ++c) {
                const double *xc = x + c * ldx + c0;
                double *yc = y + c * ldy + r0;
                for (i64 j = 0; j < tl->ncols_local; ++j) {
                    double xj = xc[j];
                    for (i64 q = tl->col_ptr[j]; q < tl->col_ptr[j + 1]; ++q)
                        yc[tl->row[q]] += tl->val[q] * xj;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------ diagonal smoothers ----- */

static double csr_get(const orc_csr *a, i64 i, i64 j, int *found) {
    i64 lo = a->row_ptr[i], hi = a->row_ptr[i + 1];
    while (lo < hi) {
        i64 mid = (lo + hi) >> 1;
        if (a->col[mid] < j) lo = mid + 1; else hi = mid;
    }
    if (lo < a->row_ptr[i + 1] && a->col[lo] == j) { *found = 1; return a->val[lo]; }
    *found = 0; return 0.0;
}

/* new_l1, smoothers.rs:63-76: d_i = 1 / sum_j |a_ij|, summed in triplet (row-major) order. */
ORC_API int orc_new_l1(const orc_csr *a, double *d) {
    for (i64 i = 0; i < a->nrows; ++i) {
        double s = 0.0;
        for (i64 q = a->row_ptr[i]; q < a->row_ptr[i + 1]; ++q) s += fabs(a->val[q]);
        d[i] = 1.0 / s;
    }
    return 0;
}
/* new_l2, smoothers.rs:43-61: d_i = 1 / sum_j |a_ij| * (sqrt(a_ii)/sqrt(a_jj)). */
ORC_API int orc_new_l2(const orc_csr *a, double *d) {
    double *ds = (double *)xmalloc(sizeof(double) * (size_t)a->nrows);
    for (i64 i = 0; i < a->nrows; ++i) {
        int f; double v = csr_get(a, i, i, &f);
        if (!f) { free(ds); return 1; } /* .unwrap() panics, smoothers.rs:46 */
        ds[i] = sqrt(v);
    }
    for (i64 i = 0; i < a->nrows; ++i) {
        double s = 0.0;
        for (i64 q = a->row_ptr[i]; q < a->row_ptr[i + 1]; ++q) {
            double scale = ds[i] / ds[a->col[q]];
            s += fabs(a->val[q]) * scale;
        }
        d[i] = 1.0 / s;
    }
    free(ds);
    return 0;
}
/* new_jacobi, smoothers.rs:78-86: d_i = omega / a_ii. */
ORC_API int orc_new_jacobi(const orc_csr *a, double omega, double *d) {
    for (i64 i = 0; i < a->nrows; ++i) {
        int f; double v = csr_get(a, i, i, &f);
        if (!f) return 1;
        d[i] = omega / v;
    }
    return 0;
}

/* ------------------------------------------------------------------ sparse x sparse -------- */

/* faer `&SparseRowMat * &SparseRowMat` [faer-recalled] as used at interpolation/mod.rs:828,
 * :938: Gustavson row-wise product; output pattern is the structural product, sorted per row,
 * NOT pruned; each entry accumulates its contributions in ascending inner index.
 * Rows are independent => OpenMP over rows does not change any value. */
ORC_API orc_csr *orc_spgemm(const orc_csr *a, const orc_csr *b) {
    if (a->ncols != b->nrows) return NULL;
    i64 n = a->nrows, nc = b->ncols;
    i64 *rp = (i64 *)xcalloc((size_t)n + 1, sizeof(i64));
#pragma omp parallel
    {
        i64 *mark = (i64 *)xmalloc(sizeof(i64) * (size_t)nc);
        for (i64 j = 0; j < nc; ++j) mark[j] = -1;
#pragma omp for schedule(static)
        for (i64 i = 0; i < n; ++i) {
            i64 c = 0;
            for (i64 q = a->row_ptr[i]; q < a->row_ptr[i + 1]; ++q) {
                i64 kk = a->col[q];
                for (i64 p = b->row_ptr[kk]; p < b->row_ptr[kk + 1]; ++p)
                    if (mark[b->col[p]] != i) { mark[b->col[p]] = i; ++c; }
            }
            rp[i + 1] = c;
        }
        free(mark);
    }
    for (i64 i = 0; i < n; ++i) rp[i + 1] += rp[i];
    orc_csr *m = csr_alloc(n, nc, rp[n]);
    memcpy(m->row_ptr, rp, sizeof(i64) * (size_t)(n + 1));
    free(rp);
#pragma omp parallel
    {
        i64 *pos = (i64 *)xmalloc(sizeof(i64) * (size_t)nc);
        for (i64 j = 0; j < nc; ++j) pos[j] = -1;
#pragma omp for schedule(static)
        for (i64 i = 0; i < n; ++i) {
            i64 base = m->row_ptr[i], c = 0;
            /* symbolic: collect then sort the column set of this row */
            for (i64 q = a->row_ptr[i]; q < a->row_ptr[i + 1]; ++q) {
                i64 kk = a->col[q];
                for (i64 p = b->row_ptr[kk]; p < b->row_ptr[kk + 1]; ++p) {
                    i64 j = b->col[p];
                    if (pos[j] < base) { pos[j] = base; m->col[base + c++] = j; }
                }
            }
            /* insertion sort (rows are short; nearly sorted input) */
            for (i64 u = base + 1; u < base + c; ++u) {
                i64 key = m->col[u], w = u - 1;
                while (w >= base && m->col[w] > key) { m->col[w + 1] = m->col[w]; --w; }
                m->col[w + 1] = key;
            }
            for (i64 u = 0; u < c; ++u) { pos[m->col[base + u]] = base + u; m->val[base + u] = 0.0; }
            /* numeric: ascending k, sequential adds */
            for (i64 q = a->row_ptr[i]; q < a->row_ptr[i + 1]; ++q) {
                i64 kk = a->col[q]; double av = a->val[q];
                for (i64 p = b->row_ptr[kk]; p < b->row_ptr[kk + 1]; ++p)
                    m->val[pos[b->col[p]]] += av * b->val[p];
            }
        }
        free(pos);
    }
    return m;
}

/* p.transpose().to_row_major(), interpolation/mod.rs:824-827: CSR of the transpose, rows sorted
 * (counting sort; source rows visited ascending). */
ORC_API orc_csr *orc_transpose(const orc_csr *a) {
    i64 nnz = a->row_ptr[a->nrows];
    orc_csr *t = csr_alloc(a->ncols, a->nrows, nnz);
    for (i64 q = 0; q < nnz; ++q) t->row_ptr[a->col[q] + 1]++;
    for (i64 j = 0; j < a->ncols; ++j) t->row_ptr[j + 1] += t->row_ptr[j];
    i64 *fill = (i64 *)xmalloc(sizeof(i64) * (size_t)(a->ncols + 1));
    memcpy(fill, t->row_ptr, sizeof(i64) * (size_t)(a->ncols + 1));
    for (i64 i = 0; i < a->nrows; ++i)
        for (i64 q = a->row_ptr[i]; q < a->row_ptr[i + 1]; ++q) {
            i64 pos = fill[a->col[q]]++;
            t->col[pos] = i; t->val[pos] = a->val[q];
        }
    free(fill);
    return t;
}

/* smooth_interpolation, interpolation/mod.rs:927-946: S = A*P; S[i,:] *= -(w * (1/a_ii));
 * S += P (add_assign requires pattern(P) subset of pattern(S), else panics).  Returns NULL when
 * a diagonal is missing / <= 1e-6 (:933-934 assert) or the pattern condition fails. */
ORC_API orc_csr *orc_smooth_interpolation(const orc_csr *a, const orc_csr *p, double w) {
    orc_csr *s = orc_spgemm(a, p);
    if (!s) return NULL;
    for (i64 i = 0; i < a->nrows; ++i) {
        int f; double dv = csr_get(a, i, i, &f);
        if (!f || !(dv > 1e-6)) { orc_csr_free(s); return NULL; }
        double scalar = w * (1.0 / dv);
        for (i64 q = s->row_ptr[i]; q < s->row_ptr[i + 1]; ++q) s->val[q] *= -scalar;
    }
    for (i64 i = 0; i < p->nrows; ++i) {
        i64 u = s->row_ptr[i];
        for (i64 q = p->row_ptr[i]; q < p->row_ptr[i + 1]; ++q) {
            while (u < s->row_ptr[i + 1] && s->col[u] < p->col[q]) ++u;
            if (u == s->row_ptr[i + 1] || s->col[u] != p->col[q]) { orc_csr_free(s); return NULL; }
            s->val[u] += p->val[q];
        }
    }
    return s;
}

/* ------------------------------------------------------------------ small dense helpers ---- */

/* Thin SVD of an m x k column-major matrix (m >= 1) by one-sided (Hestenes) Jacobi.
 * Stands in for faer `thin_svd` (interpolation/mod.rs:770) whose column-sign convention is
 * unverifiable.  Convention here: singular values descending; for each triplet
 * the largest-|.| entry of v_j is made positive.  u: m x k, s: k, v: k x k (column-major).
 * Columns of u with s_j == 0 are left zero (rank-deficient aggregates). */
static void thin_svd(i64 m, i64 k, const double *a, double *u, double *s, double *v) {
    memcpy(u, a, sizeof(double) * (size_t)(m * k));
    for (i64 i = 0; i < k * k; ++i) v[i] = 0.0;
    for (i64 i = 0; i < k; ++i) v[i + i * k] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (i64 p = 0; p < k - 1; ++p)
            for (i64 q = p + 1; q < k; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (i64 i = 0; i < m; ++i) {
                    alpha += u[i + p * m] * u[i + p * m];
                    beta += u[i + q * m] * u[i + q * m];
                    gamma += u[i + p * m] * u[i + q * m];
                }
                if (gamma == 0.0) continue;
                double lim = sqrt(alpha * beta);
                if (fabs(gamma) <= 1e-300 || fabs(gamma) <= 1e-16 * lim) continue;
                off = fmax(off, fabs(gamma) / (lim > 0 ? lim : 1.0));
                double zeta = (beta - alpha) / (2.0 * gamma);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
                for (i64 i = 0; i < m; ++i) {
                    double up = u[i + p * m], uq = u[i + q * m];
                    u[i + p * m] = c * up - sn * uq;
                    u[i + q * m] = sn * up + c * uq;
                }
                for (i64 i = 0; i < k; ++i) {
                    double vp = v[i + p * k], vq = v[i + q * k];
                    v[i + p * k] = c * vp - sn * vq;
                    v[i + q * k] = sn * vp + c * vq;
                }
            }
        if (off < 1e-15) break;
    }
    for (i64 j = 0; j < k; ++j) {
        double nn = 0;
        for (i64 i = 0; i < m; ++i) nn += u[i + j * m] * u[i + j * m];
        s[j] = sqrt(nn);
    }
    /* sort descending (selection sort, k is tiny) */
    for (i64 j = 0; j < k; ++j) {
        i64 best = j;
        for (i64 l = j + 1; l < k; ++l) if (s[l] > s[best]) best = l;
        if (best != j) {
            double ts = s[j]; s[j] = s[best]; s[best] = ts;
            for (i64 i = 0; i < m; ++i) { double t = u[i + j * m]; u[i + j * m] = u[i + best * m]; u[i + best * m] = t; }
            for (i64 i = 0; i < k; ++i) { double t = v[i + j * k]; v[i + j * k] = v[i + best * k]; v[i + best * k] = t; }
        }
    }
    for (i64 j = 0; j < k; ++j) {
        if (s[j] > 0) for (i64 i = 0; i < m; ++i) u[i + j * m] /= s[j];
        else for (i64 i = 0; i < m; ++i) u[i + j * m] = 0.0;
        i64 big = 0;
        for (i64 i = 1; i < k; ++i) if (fabs(v[i + j * k]) > fabs(v[big + j * k])) big = i;
        if (v[big + j * k] < 0) {
            for (i64 i = 0; i < k; ++i) v[i + j * k] = -v[i + j * k];
            for (i64 i = 0; i < m; ++i) u[i + j * m] = -u[i + j * m];
        }
    }
}

/* Thin Q of an n x k column-major matrix (`coarse_nn.qr().compute_thin_Q()`, hierarchy.rs:228)
 * by modified Gram-Schmidt applied twice; R has a positive diagonal (faer's Householder sign
 * is unverifiable; for k = 1 this is v/||v||). In place. */
ORC_API void orc_thin_q(i64 n, i64 k, double *a, i64 lda) {
    for (i64 j = 0; j < k; ++j) {
        double *aj = a + j * lda;
        for (int pass = 0; pass < 2; ++pass)
            for (i64 l = 0; l < j; ++l) {
                const double *al = a + l * lda;
                double d = 0;
                for (i64 i = 0; i < n; ++i) d += al[i] * aj[i];
                for (i64 i = 0; i < n; ++i) aj[i] -= d * al[i];
            }
        double nn = 0;
        for (i64 i = 0; i < n; ++i) nn += aj[i] * aj[i];
        nn = sqrt(nn);
        for (i64 i = 0; i < n; ++i) aj[i] /= nn;
    }
}

/* Tentative prolongator of smoothed_aggregation, interpolation/mod.rs:747-809.
 * aggregates given CSR-style (agg_ptr[n_aggs+1], agg_nodes ascending inside each aggregate =
 * BTreeSet iteration order, partitioners/mod.rs:23-27).  near_null: n_fine x k column-major.
 * Outputs P (n_fine x n_aggs*cand) and coarse_nn (n_aggs*cand x k, column-major, ld = n_aggs*cand).
 * Returns NULL if an aggregate is smaller than cand (assert :757-762). */
ORC_API orc_csr *orc_tentative_p(i64 n_fine, i64 block_size, i64 k, i64 cand, const double *near_null,
                                 i64 ldnn, i64 n_aggs, const i64 *agg_ptr, const i64 *agg_nodes,
                                 double *coarse_nn) {
    i64 nc = n_aggs * cand;
    i64 total_nodes = agg_ptr[n_aggs];
    i64 nt = total_nodes * block_size * cand;
    i64 *tr = (i64 *)xmalloc(sizeof(i64) * (size_t)nt), *tc = (i64 *)xmalloc(sizeof(i64) * (size_t)nt);
    double *tv = (double *)xmalloc(sizeof(double) * (size_t)nt);
    i64 t = 0; int bad = 0;
    for (i64 ci = 0; ci < n_aggs && !bad; ++ci) {
        i64 na = agg_ptr[ci + 1] - agg_ptr[ci];
        i64 m = na * block_size;
        if (m < cand || cand > k) { bad = 1; break; }
        double *local = (double *)xmalloc(sizeof(double) * (size_t)(m * k));
        double *u = (double *)xmalloc(sizeof(double) * (size_t)(m * k));
        double *s = (double *)xmalloc(sizeof(double) * (size_t)k);
        double *v = (double *)xmalloc(sizeof(double) * (size_t)(k * k));
        for (i64 lj = 0; lj < na; ++lj) {
            i64 j = agg_nodes[agg_ptr[ci] + lj];
            for (i64 o = 0; o < block_size; ++o)
                for (i64 c = 0; c < k; ++c)
                    local[lj * block_size + o + c * m] = near_null[j * block_size + o + c * ldnn];
        }
        thin_svd(m, k, local, u, s, v);
        /* r = S * V^T ; coarse_nn rows = first cand rows of r  (:788-792) */
        for (i64 r = 0; r < cand; ++r)
            for (i64 c = 0; c < k; ++c) coarse_nn[ci * cand + r + c * nc] = s[r] * v[c + r * k];
        for (i64 li = 0; li < na; ++li) {
            i64 fi = agg_nodes[agg_ptr[ci] + li];
            for (i64 oi = 0; oi < block_size; ++oi)
                for (i64 oj = 0; oj < cand; ++oj) {
                    tr[t] = fi * block_size + oi; tc[t] = ci * cand + oj;
                    tv[t++] = u[li * block_size + oi + oj * m];
                }
        }
        free(local); free(u); free(s); free(v);
    }
    orc_csr *p = bad ? NULL : orc_csr_from_triplets(n_fine, nc, t, tr, tc, tv);
    free(tr); free(tc); free(tv);
    return p;
}

/* Dense Cholesky factor + solve: stands in for SparseCholeskySolve (coarse_solvers.rs:173-206);
 * a sparse and a dense LLT solve the same system and differ by rounding only. */
typedef struct { i64 n; double *l; } orc_llt;
ORC_API orc_llt *orc_llt_new(const orc_csr *a) {
    i64 n = a->nrows;
    orc_llt *f = (orc_llt *)xmalloc(sizeof(orc_llt));
    f->n = n; f->l = (double *)xcalloc((size_t)(n * n), sizeof(double));
    double *l = f->l; /* column-major lower */
    for (i64 i = 0; i < n; ++i)
        for (i64 q = a->row_ptr[i]; q < a->row_ptr[i + 1]; ++q)
            if (a->col[q] <= i) l[i + a->col[q] * n] = a->val[q];
    for (i64 j = 0; j < n; ++j) {
        double d = l[j + j * n];
        for (i64 kk = 0; kk < j; ++kk) d -= l[j + kk * n] * l[j + kk * n];
        if (!(d > 0)) { free(f->l); free(f); return NULL; }
        d = sqrt(d); l[j + j * n] = d;
        for (i64 i = j + 1; i < n; ++i) {
            double sv = l[i + j * n];
            for (i64 kk = 0; kk < j; ++kk) sv -= l[i + kk * n] * l[j + kk * n];
            l[i + j * n] = sv / d;
        }
    }
    return f;
}
ORC_API void orc_llt_free(orc_llt *f) { if (f) { free(f->l); free(f); } }
ORC_API void orc_llt_solve_in_place(const orc_llt *f, double *x, i64 ldx, i64 k) {
    i64 n = f->n; const double *l = f->l;
    for (i64 c = 0; c < k; ++c) {
        double *xc = x + c * ldx;
        for (i64 i = 0; i < n; ++i) {
            double sv = xc[i];
            for (i64 j = 0; j < i; ++j) sv -= l[i + j * n] * xc[j];
            xc[i] = sv / l[i + i * n];
        }
        for (i64 i = n - 1; i >= 0; --i) {
            double sv = xc[i];
            for (i64 j = i + 1; j < n; ++j) sv -= l[j + i * n] * xc[j];
            xc[i] = sv / l[i + i * n];
        }
    }
}

/* ------------------------------------------------------------------ Multigrid -------------- */

enum { ORC_SM_DIAG = 0, ORC_SM_LLT = 1, ORC_SM_BLOCK = 2 };

typedef struct {
    i64 n_aggs;
    i64 *agg_ptr, *agg_nodes; /* copies */
    orc_llt **blocks;
} orc_blocksm;

typedef struct {
    const orc_csr *a;          /* borrowed */
    orc_parspmm *par;          /* owned, may be NULL */
    int sm_kind;
    double *d;                 /* owned copy (DIAG) */
    orc_llt *llt;              /* owned (LLT) */
    orc_blocksm *blk;          /* owned (BLOCK) */
    const orc_csr *r, *p;      /* borrowed; NULL on level 0 */
} orc_level;

typedef struct {
    int nlevels, cap;
    orc_level *lv;
    int mu, nu;
} orc_mg;

static void level_apply_a(const orc_level *l, const double *x, i64 ldx, double *y, i64 ldy, i64 k) {
    if (l->par) orc_parspmm_apply(l->par, x, ldx, y, ldy, k);
    else orc_spmm_csr(l->a, x, ldx, y, ldy, k);
}

/* diagonally_compensate, block_smoothers.rs:293-324 (scalar case): principal sub-matrix on the
 * aggregate; every coupling a_ij to a node j outside the aggregate becomes an extra (ic,ic)
 * triplet 0.5*sqrt(a_ii/a_jj)*|a_ij| pushed in row order, so the diagonal is the in-order sum
 * formed by try_new_from_triplets.  `nodes` ascending (BTreeSet order). */
static orc_csr *diag_compensate(const orc_csr *a, const i64 *nodes, i64 na, const double *diag) {
    i64 cap = 0;
    for (i64 li = 0; li < na; ++li) cap += a->row_ptr[nodes[li] + 1] - a->row_ptr[nodes[li]];
    i64 *tr = (i64 *)xmalloc(sizeof(i64) * (size_t)cap), *tc = (i64 *)xmalloc(sizeof(i64) * (size_t)cap);
    double *tv = (double *)xmalloc(sizeof(double) * (size_t)cap);
    i64 t = 0;
    for (i64 li = 0; li < na; ++li) {
        i64 i = nodes[li];
        for (i64 p = a->row_ptr[i]; p < a->row_ptr[i + 1]; ++p) {
            i64 j = a->col[p];
            i64 lo = 0, hi = na;
            while (lo < hi) { i64 mid = (lo + hi) >> 1; if (nodes[mid] < j) lo = mid + 1; else hi = mid; }
            if (lo < na && nodes[lo] == j) { tr[t] = li; tc[t] = lo; tv[t++] = a->val[p]; }
            else { tr[t] = li; tc[t] = li; tv[t++] = 0.5 * sqrt(diag[i] / diag[j]) * fabs(a->val[p]); }
        }
    }
    orc_csr *m = orc_csr_from_triplets(na, na, t, tr, tc, tv);
    free(tr); free(tc); free(tv);
    return m;
}

ORC_API orc_mg *orc_mg_new(void) {
    orc_mg *mg = (orc_mg *)xcalloc(1, sizeof(orc_mg));
    mg->cap = 8; mg->lv = (orc_level *)xcalloc((size_t)mg->cap, sizeof(orc_level));
    mg->mu = 1; mg->nu = 1;
    return mg;
}
/* Multigrid::new / add_level, multigrid.rs:190-239.  sm_kind DIAG: d = diagonal (copied);
 * LLT: exact Cholesky solve of `a` (SparseCholeskySolve); BLOCK: BlockSmoother over aggregates
 * (agg_ptr/agg_nodes).  nthreads > 1 wraps `a` in ParSpmmOp as MultigridConfig::build does
 * (multigrid.rs:134-148). r/p NULL for the finest level. Returns level index or -1. */
ORC_API int orc_mg_add_level(orc_mg *mg, const orc_csr *a, int sm_kind, const double *d,
                             i64 n_aggs, const i64 *agg_ptr, const i64 *agg_nodes,
                             const orc_csr *r, const orc_csr *p, int nthreads) {
    if (mg->nlevels == mg->cap) {
        mg->cap *= 2;
        mg->lv = (orc_level *)realloc(mg->lv, sizeof(orc_level) * (size_t)mg->cap);
    }
    orc_level *l = &mg->lv[mg->nlevels];
    memset(l, 0, sizeof(*l));
    l->a = a; l->r = r; l->p = p; l->sm_kind = sm_kind;
    if (mg->nlevels > 0) {
        const orc_level *prev = &mg->lv[mg->nlevels - 1];
        if (!r || !p || r->ncols != prev->a->nrows || p->nrows != prev->a->nrows ||
            r->nrows != a->nrows || p->ncols != a->nrows) return -1;
    }
    if (sm_kind == ORC_SM_DIAG) {
        l->d = (double *)xmalloc(sizeof(double) * (size_t)a->nrows);
        memcpy(l->d, d, sizeof(double) * (size_t)a->nrows);
    } else if (sm_kind == ORC_SM_LLT) {
        l->llt = orc_llt_new(a);
        if (!l->llt) return -1;
    } else if (sm_kind == ORC_SM_BLOCK) {
        orc_blocksm *b = (orc_blocksm *)xcalloc(1, sizeof(orc_blocksm));
        b->n_aggs = n_aggs;
        b->agg_ptr = (i64 *)xmalloc(sizeof(i64) * (size_t)(n_aggs + 1));
        memcpy(b->agg_ptr, agg_ptr, sizeof(i64) * (size_t)(n_aggs + 1));
        b->agg_nodes = (i64 *)xmalloc(sizeof(i64) * (size_t)agg_ptr[n_aggs]);
        memcpy(b->agg_nodes, agg_nodes, sizeof(i64) * (size_t)agg_ptr[n_aggs]);
        b->blocks = (orc_llt **)xcalloc((size_t)n_aggs, sizeof(orc_llt *));
        double *diag = (double *)xmalloc(sizeof(double) * (size_t)a->nrows);
        for (i64 i = 0; i < a->nrows; ++i) { int f; diag[i] = csr_get(a, i, i, &f); }
        int fail = 0;
#pragma omp parallel for schedule(dynamic, 64)
        for (i64 g = 0; g < n_aggs; ++g) {
            orc_csr *loc = diag_compensate(a, b->agg_nodes + b->agg_ptr[g], b->agg_ptr[g + 1] - b->agg_ptr[g], diag);
            b->blocks[g] = orc_llt_new(loc);
            if (!b->blocks[g]) fail = 1;
            orc_csr_free(loc);
        }
        free(diag);
        l->blk = b;
        if (fail) return -1;
    }
    if (nthreads > 1) l->par = orc_parspmm_new(a, nthreads);
    return mg->nlevels++;
}
ORC_API void orc_mg_set_cycle(orc_mg *mg, int mu, int nu) { mg->mu = mu; mg->nu = nu; }
ORC_API void orc_mg_free(orc_mg *mg) {
    if (!mg) return;
    for (int i = 0; i < mg->nlevels; ++i) {
        orc_level *l = &mg->lv[i];
        free(l->d); orc_llt_free(l->llt); orc_parspmm_free(l->par);
        if (l->blk) {
            for (i64 g = 0; g < l->blk->n_aggs; ++g) orc_llt_free(l->blk->blocks[g]);
            free(l->blk->blocks); free(l->blk->agg_ptr); free(l->blk->agg_nodes); free(l->blk);
        }
    }
    free(mg->lv); free(mg);
}

/* smoother.apply_in_place(r): Diag -> row scaling [faer-recalled]; SparseCholeskySolve ->
 * exact solve (coarse_solvers.rs:254-258); BlockSmoother::apply -> gather / per-block solve /
 * scatter (block_smoothers.rs:165-214). */
static void smoother_apply_in_place(const orc_level *l, double *r, i64 ld, i64 k) {
    i64 n = l->a->nrows;
    if (l->sm_kind == ORC_SM_DIAG) {
        for (i64 c = 0; c < k; ++c)
            for (i64 i = 0; i < n; ++i) r[i + c * ld] = l->d[i] * r[i + c * ld];
    } else if (l->sm_kind == ORC_SM_LLT) {
        orc_llt_solve_in_place(l->llt, r, ld, k);
    } else {
        const orc_blocksm *b = l->blk;
#pragma omp parallel for schedule(dynamic, 64)
        for (i64 g = 0; g < b->n_aggs; ++g) {
            i64 na = b->agg_ptr[g + 1] - b->agg_ptr[g];
            const i64 *nodes = b->agg_nodes + b->agg_ptr[g];
            double *part = (double *)xmalloc(sizeof(double) * (size_t)(na * k));
            for (i64 c = 0; c < k; ++c)
                for (i64 li = 0; li < na; ++li) part[li + c * na] = r[nodes[li] + c * ld];
            orc_llt_solve_in_place(b->blocks[g], part, na, k);
            for (i64 c = 0; c < k; ++c)
                for (i64 li = 0; li < na; ++li) r[nodes[li] + c * ld] = part[li + c * na];
            free(part);
        }
    }
}

/* smooth(), multigrid.rs:407-424: nu x { work = A x; r = b - work; r <- M^-1 r; x += r }.
 * Allocates like the reference (Mat::zeros per call / `b - &work` per sweep). */
static void mg_smooth(const orc_level *l, double *x, const double *b, i64 n, i64 k, int iters) {
    double *work = (double *)xcalloc((size_t)(n * k), sizeof(double));
    for (int it = 0; it < iters; ++it) {
        level_apply_a(l, x, n, work, n, k);
        double *r = (double *)xmalloc(sizeof(double) * (size_t)(n * k));
        for (i64 i = 0; i < n * k; ++i) r[i] = b[i] - work[i];
        smoother_apply_in_place(l, r, n, k);
        for (i64 i = 0; i < n * k; ++i) x[i] += r[i];
        free(r);
    }
    free(work);
}

/* Multigrid::cycle, multigrid.rs:269-380 (DEBUG branches are compile-time dead, :181). */
static void mg_cycle(const orc_mg *mg, double *v, const double *f, int level, i64 k) {
    const orc_level *l = &mg->lv[level];
    i64 n = l->a->nrows;
    double *work = (double *)xcalloc((size_t)(n * k), sizeof(double)); /* :278 */
    if (level == mg->nlevels - 1) {
        /* smoother.apply(v, f) (:292): out-of-place M^-1 f */
        memcpy(v, f, sizeof(double) * (size_t)(n * k));
        smoother_apply_in_place(l, v, n, k);
        free(work);
        return;
    }
    mg_smooth(l, v, f, n, k, mg->nu);                                   /* :314-322 */
    const orc_level *lc = &mg->lv[level + 1];
    i64 nc = lc->a->nrows;
    double *vc = (double *)xcalloc((size_t)(nc * k), sizeof(double));   /* :337 */
    double *fc = (double *)xcalloc((size_t)(nc * k), sizeof(double));   /* :338 */
    level_apply_a(l, v, n, work, n, k);                                 /* :341 */
    for (i64 i = 0; i < n * k; ++i) work[i] = f[i] - work[i];           /* :342 */
    orc_spmm_csr(lc->r, work, n, fc, nc, k);                            /* :343 */
    for (int m = 0; m < mg->mu; ++m) mg_cycle(mg, vc, fc, level + 1, k);/* :345-347 */
    orc_spmm_csr(lc->p, vc, nc, work, n, k);                            /* :349 */
    for (i64 i = 0; i < n * k; ++i) v[i] += work[i];                    /* :350 */
    mg_smooth(l, v, f, n, k, mg->nu);                                   /* :361-369 */
    free(vc); free(fc); free(work);
}

/* Multigrid::apply -> init_cycle, multigrid.rs:469-473, 251-267: out = 0; v = 0;
 * cycle(v, rhs, 0); out += v.  out/rhs: n x k column-major, ld = n. */
ORC_API void orc_mg_apply(const orc_mg *mg, double *out, const double *rhs, i64 k) {
    i64 n = mg->lv[0].a->nrows;
    for (i64 i = 0; i < n * k; ++i) out[i] = 0.0;
    double *v = (double *)xcalloc((size_t)(n * k), sizeof(double));
    mg_cycle(mg, v, rhs, 0, k);
    for (i64 i = 0; i < n * k; ++i) out[i] += v[i];
    free(v);
}

/* One smoothing call on a single level, exposed for kernel-parity tests of the fused sweep. */
ORC_API void orc_smooth_diag(const orc_csr *a, const double *d, double *x, const double *b, i64 k,
                             int iters) {
    orc_level l; memset(&l, 0, sizeof(l));
    l.a = a; l.sm_kind = ORC_SM_DIAG; l.d = (double *)d;
    mg_smooth(&l, x, b, a->nrows, k, iters);
}

/* BlockSmoother stand-alone apply (block_smoothers.rs:165-214) for tier-2 parity tests. */
ORC_API int orc_block_smoother_apply(const orc_csr *a, i64 n_aggs, const i64 *agg_ptr,
                                     const i64 *agg_nodes, double *r, i64 k) {
    orc_mg *mg = orc_mg_new();
    int rc = orc_mg_add_level(mg, a, ORC_SM_BLOCK, NULL, n_aggs, agg_ptr, agg_nodes, NULL, NULL, 1);
    if (rc >= 0) smoother_apply_in_place(&mg->lv[0], r, a->nrows, k);
    orc_mg_free(mg);
    return rc < 0;
}

/* StationaryIteration::apply, smoothers.rs:146-159, with a Diag preconditioner (as used on the
 * coarse near-null, hierarchy.rs:219-226):  x = M rhs; repeat iters-1: r = A x; r = x - r (sic);
 * out = M r; x += out.  In place on `io` (n x k, ld = n) like apply_in_place. */
ORC_API void orc_stationary_iteration(const orc_csr *a, const double *d, int iters, double *io, i64 k) {
    i64 n = a->nrows;
    double *x = (double *)xmalloc(sizeof(double) * (size_t)(n * k));
    double *r = (double *)xmalloc(sizeof(double) * (size_t)(n * k));
    double *o = (double *)xmalloc(sizeof(double) * (size_t)(n * k));
    for (i64 c = 0; c < k; ++c) for (i64 i = 0; i < n; ++i) x[i + c * n] = d[i] * io[i + c * n];
    for (int it = 1; it < iters; ++it) {
        orc_spmm_csr(a, x, n, r, n, k);
        for (i64 i = 0; i < n * k; ++i) r[i] = x[i] - r[i];
        for (i64 c = 0; c < k; ++c) for (i64 i = 0; i < n; ++i) o[i + c * n] = d[i] * r[i + c * n];
        for (i64 i = 0; i < n * k; ++i) x[i] += o[i];
    }
    memcpy(io, x, sizeof(double) * (size_t)(n * k));
    free(x); free(r); free(o);
}

/* ------------------------------------------------------------------ solvers ---------------- */

static double nrm2(const double *x, i64 n) {
    double s = 0;
    for (i64 i = 0; i < n; ++i) s += x[i] * x[i];
    return sqrt(s);
}
static double dot(const double *x, const double *y, i64 n) {
    double s = 0;
    for (i64 i = 0; i < n; ++i) s += x[i] * y[i];
    return s;
}

enum { ORC_PC_NONE = 0, ORC_PC_DIAG = 1, ORC_PC_MG = 2 };
typedef struct { i64 iters; double abs_res, rel_res; int status; } orc_cg_info;

static void pc_apply(int kind, const double *d, const orc_mg *mg, double *z, const double *r, i64 n) {
    if (kind == ORC_PC_MG) orc_mg_apply(mg, z, r, 1);
    else if (kind == ORC_PC_DIAG) for (i64 i = 0; i < n; ++i) z[i] = d[i] * r[i];
    else memcpy(z, r, sizeof(double) * (size_t)n);
}

/* faer::matrix_free::conjugate_gradient [faer-recalled] as driven by utils.rs:574-609 and
 * simple_geometric.rs:229-267, single right-hand side.  threshold = max(abs_tol, rel_tol*||b||);
 * zero ||b|| => x = 0, 0 iterations; r = b - A x0 when the guess may be non-zero, else r = b;
 * converged before the loop if ||r|| < threshold; z = M r; p = z; loop: q = A p;
 * alpha = (r.z)/(p.q); x += alpha p; r -= alpha q; iter_count = it+1 if ||r|| < threshold;
 * z = M r; beta = (r.z)new/(r.z)old; p = z + beta p.  status: 0 ok, 1 NoConvergence,
 * 2 non-positive-definite operator/preconditioner.  `par` (nullable) is the ParSpmmOp for A. */
ORC_API int orc_pcg(const orc_csr *a, const orc_parspmm *par, int pc_kind, const double *d,
                    const orc_mg *mg, const double *b, double *x, double rel_tol, double abs_tol,
                    i64 max_iters, int zero_guess, orc_cg_info *info) {
    i64 n = a->nrows;
    double b_norm = nrm2(b, n);
    info->iters = 0; info->abs_res = 0; info->rel_res = 0; info->status = 0;
    if (b_norm == 0.0) { memset(x, 0, sizeof(double) * (size_t)n); return 0; }
    double thr = rel_tol * b_norm; if (abs_tol > thr) thr = abs_tol;
    double *r = (double *)xmalloc(sizeof(double) * (size_t)n), *p = (double *)xmalloc(sizeof(double) * (size_t)n);
    double *z = (double *)xmalloc(sizeof(double) * (size_t)n), *q = (double *)xmalloc(sizeof(double) * (size_t)n);
    if (zero_guess) { memset(x, 0, sizeof(double) * (size_t)n); memcpy(r, b, sizeof(double) * (size_t)n); }
    else {
        if (par) orc_parspmm_apply(par, x, n, r, n, 1); else orc_spmm_csr(a, x, n, r, n, 1);
        for (i64 i = 0; i < n; ++i) r[i] = b[i] - r[i];
    }
    double rn = nrm2(r, n);
    int rc = 1;
    if (rn < thr) { rc = 0; goto done; }
    pc_apply(pc_kind, d, mg, z, r, n);
    memcpy(p, z, sizeof(double) * (size_t)n);
    double rtz = dot(r, z, n);
    for (i64 it = 0; it < max_iters; ++it) {
        if (par) orc_parspmm_apply(par, p, n, q, n, 1); else orc_spmm_csr(a, p, n, q, n, 1);
        double ptq = dot(p, q, n);
        if (!(ptq > 0.0)) { rc = 2; break; }
        double alpha = rtz / ptq;
        for (i64 i = 0; i < n; ++i) x[i] += alpha * p[i];
        for (i64 i = 0; i < n; ++i) r[i] -= alpha * q[i];
        rn = nrm2(r, n);
        info->iters = it + 1;
        if (rn < thr) { rc = 0; break; }
        pc_apply(pc_kind, d, mg, z, r, n);
        double rtz_new = dot(r, z, n);
        if (!(rtz_new > 0.0)) { rc = 2; break; }
        double beta = rtz_new / rtz;
        rtz = rtz_new;
        for (i64 i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
    }
done:
    info->abs_res = rn; info->rel_res = rn / b_norm; info->status = rc;
    free(r); free(p); free(z); free(q);
    return rc;
}

/* stationary_solver, examples/simple_geometric.rs:117-158: loop { r = b - A x; iter += 1;
 * stop if ||r||/||b|| < tol or iter >= max_iter; r <- M r; x += r }.  Returns iteration count. */
ORC_API i64 orc_stationary_solver(const orc_csr *a, int pc_kind, const double *d, const orc_mg *mg,
                                  const double *b, double *x, i64 max_iter, double rel_tol) {
    i64 n = a->nrows, iter = 0;
    double *work = (double *)xmalloc(sizeof(double) * (size_t)n), *r = (double *)xmalloc(sizeof(double) * (size_t)n);
    double *z = (double *)xmalloc(sizeof(double) * (size_t)n);
    double b_norm = nrm2(b, n);
    for (;;) {
        orc_spmm_csr(a, x, n, work, n, 1);
        for (i64 i = 0; i < n; ++i) r[i] = b[i] - work[i];
        double rel = nrm2(r, n) / b_norm;
        ++iter;
        if (rel < rel_tol || iter >= max_iter) break;
        pc_apply(pc_kind, d, mg, z, r, n);
        for (i64 i = 0; i < n; ++i) x[i] += z[i];
    }
    free(work); free(r); free(z);
    return iter;
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
