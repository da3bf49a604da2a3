// smoothers.cu -- level smoothers: Diag (L1 / L2 / Jacobi), exact coarse solve, block smoother.
// Reference: src/preconditioners/smoothers.rs:43-86, coarse_solvers.rs:173-206,
// block_smoothers.rs:88-146,293-324, and smooth() in multigrid.rs:407-424.
#include <cmath>

#include "common.cuh"

namespace famg {

// ---------------------------------------------------------------- Diag setup (one pass over A)
// Thread per row, entries visited in ascending column order == triplet_iter order, so the sums
// round exactly like the reference's sequential loops.
__global__ void diag_l1_kernel(const int *__restrict__ row_ptr, const double *__restrict__ val, int n, double *__restrict__ d) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int q = row_ptr[i]; q < row_ptr[i + 1]; ++q) s += fabs(val[q]);
    d[i] = 1.0 / s;  // d.recip()  (smoothers.rs:72-75)
}

// extracts a_ii (binary search like mat.get(i,i)); flags missing diagonals
__global__ void diag_extract_kernel(const int *__restrict__ row_ptr, const int *__restrict__ col, const double *__restrict__ val,
                                    int n, double *__restrict__ diag, int *__restrict__ missing) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = row_ptr[i], hi = row_ptr[i + 1];
    const int end = hi;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (col[mid] < i) lo = mid + 1; else hi = mid;
    }
    if (lo < end && col[lo] == i) diag[i] = val[lo];
    else { diag[i] = 0.0; atomicAdd(missing, 1); }
}

__global__ void diag_sqrt_kernel(double *__restrict__ v, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = sqrt(v[i]);
}

__global__ void diag_l2_kernel(const int *__restrict__ row_ptr, const int *__restrict__ col, const double *__restrict__ val, int n,
                               const double *__restrict__ dsqrt, double *__restrict__ d) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    const double di = dsqrt[i];
    for (int q = row_ptr[i]; q < row_ptr[i + 1]; ++q) {
        const double scale = di / dsqrt[col[q]];
        s += fabs(val[q]) * scale;
    }
    d[i] = 1.0 / s;
}

__global__ void diag_jacobi_kernel(const double *diag, int n, double omega, double *d) {  // diag may alias d
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = omega / diag[i];
}

// a_ii by linear search (row slabs of a distributed operator are not sorted by local column id)
__global__ void diag_extract_linear_kernel(const int *__restrict__ row_ptr, const int *__restrict__ col, const double *__restrict__ val,
                                           int n, double *__restrict__ diag, int *__restrict__ missing) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = 0.0; bool found = false;
    for (int q = row_ptr[i]; q < row_ptr[i + 1]; ++q) if (col[q] == i) { v = val[q]; found = true; break; }
    diag[i] = v;
    if (!found) atomicAdd(missing, 1);
}

famg_status diag_from_rows(const famg_csr *a, int kind, double omega, double *d_out) {
    famg_ctx *ctx = a->ctx;
    const int n = (int)a->nrows;
    if (n == 0) return FAMG_OK;
    const unsigned grid = (unsigned)ceil_div(n, 256);
    if (kind == FAMG_DIAG_L1) {
        diag_l1_kernel<<<grid, 256, 0, ctx->stream>>>(a->row_ptr, a->val, n, d_out);
        count_launch(ctx);
        KERNEL_CHECK();
        return FAMG_OK;
    }
    if (kind != FAMG_DIAG_JACOBI) FAMG_FAIL(FAMG_ERR_UNSUPPORTED, "distributed levels support the L1 and Jacobi diagonal smoothers");
    int *missing = nullptr;
    FAMG_TRY(pool_alloc(ctx, 256, (void **)&missing));
    cudaMemsetAsync(missing, 0, sizeof(int), ctx->stream);
    diag_extract_linear_kernel<<<grid, 256, 0, ctx->stream>>>(a->row_ptr, a->col, a->val, n, d_out, missing);
    diag_jacobi_kernel<<<grid, 256, 0, ctx->stream>>>(d_out, n, omega, d_out);
    count_launch(ctx, 2);
    int h_missing = 0;
    cudaError_t e = cudaMemcpyAsync(&h_missing, missing, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    pool_free(ctx, missing, 256);
    if (e != cudaSuccess) FAMG_FAIL(FAMG_ERR_CUDA, "diag setup failed: %s", cudaGetErrorString(e));
    if (h_missing) FAMG_FAIL(FAMG_ERR_NUMERIC, "matrix has %d rows without a diagonal entry", h_missing);
    return FAMG_OK;
}

void smoother_release(famg_smoother *s) {
    if (!s) return;
    if (s->refs.fetch_sub(1) == 1) {
        cudaFree(s->d); cudaFree(s->inv);
        if (s->minv) csr_release(s->minv);
        delete s;
    }
}

famg_status smoother_apply_dev(const famg_smoother *s, const double *in, int64_t ldi, double *out, int64_t ldo, int k,
                               cudaStream_t st) {
    famg_ctx *ctx = s->ctx;
    if (s->kind == SM_DIAG) return vec_scale_rows(ctx, s->d, in, ldi, out, ldo, s->n, k, st);
    if (s->kind == SM_DENSE_INV) return dense_gemv(ctx, s->inv, s->n, in, ldi, out, ldo, k, st);
    if (in == out) FAMG_FAIL(FAMG_ERR_INVALID, "block smoother apply cannot run in place");
    SpmvArgs a; a.a = s->minv; a.epi = EPI_SPMV; a.x = in; a.ldx = ldi; a.y = out; a.ldy = ldo; a.k = k; a.stream = st;
    return spmv_launch(a);
}

}  // namespace famg

using namespace famg;

extern "C" {

famg_status famg_smoother_diag(const famg_csr *a, int kind, double omega, famg_smoother **out) {
    if (!a || !out) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    *out = nullptr;
    if (a->nrows != a->ncols) FAMG_FAIL(FAMG_ERR_INVALID, "diagonal smoother needs a square matrix");
    famg_ctx *ctx = a->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int n = (int)a->nrows;
    famg_smoother *s = new famg_smoother();
    s->ctx = ctx; s->kind = SM_DIAG; s->n = n;
    famg_status st = dev_alloc(&s->d, n);
    if (st != FAMG_OK) { smoother_release(s); return st; }
    const unsigned grid = (unsigned)ceil_div(std::max(n, 1), 256);
    if (kind == FAMG_DIAG_L1) {
        diag_l1_kernel<<<grid, 256, 0, ctx->stream>>>(a->row_ptr, a->val, n, s->d);
        count_launch(ctx);
    } else if (kind == FAMG_DIAG_L2 || kind == FAMG_DIAG_JACOBI) {
        double *diag = nullptr; int *missing = nullptr;
        st = dev_alloc(&diag, n);
        if (st == FAMG_OK) st = dev_alloc(&missing, 1);
        if (st != FAMG_OK) { cudaFree(diag); smoother_release(s); return st; }
        cudaMemsetAsync(missing, 0, sizeof(int), ctx->stream);
        diag_extract_kernel<<<grid, 256, 0, ctx->stream>>>(a->row_ptr, a->col, a->val, n, diag, missing);
        if (kind == FAMG_DIAG_L2) {
            diag_sqrt_kernel<<<grid, 256, 0, ctx->stream>>>(diag, n);
            diag_l2_kernel<<<grid, 256, 0, ctx->stream>>>(a->row_ptr, a->col, a->val, n, diag, s->d);
            count_launch(ctx, 3);
        } else {
            diag_jacobi_kernel<<<grid, 256, 0, ctx->stream>>>(diag, n, omega, s->d);
            count_launch(ctx, 2);
        }
        int h_missing = 0;
        cudaError_t e = cudaMemcpyAsync(&h_missing, missing, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        cudaFree(diag); cudaFree(missing);
        if (e != cudaSuccess) { smoother_release(s); FAMG_FAIL(FAMG_ERR_CUDA, "diag setup failed: %s", cudaGetErrorString(e)); }
        if (h_missing) {  // mat.get(i,i).unwrap() panics in the reference (smoothers.rs:46,83)
            smoother_release(s);
            FAMG_FAIL(FAMG_ERR_NUMERIC, "%d rows have no diagonal entry", h_missing);
        }
    } else {
        smoother_release(s);
        FAMG_FAIL(FAMG_ERR_INVALID, "unknown diagonal smoother kind %d", kind);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { smoother_release(s); FAMG_FAIL(FAMG_ERR_CUDA, "diag setup failed: %s", cudaGetErrorString(e)); }
    *out = s;
    return FAMG_OK;
}

famg_status famg_smoother_diag_from_host(famg_ctx *ctx, int64_t n, const double *d, famg_smoother **out) {
    if (!ctx || !out || (n && !d)) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    *out = nullptr;
    CUDA_TRY(cudaSetDevice(ctx->device));
    famg_smoother *s = new famg_smoother();
    s->ctx = ctx; s->kind = SM_DIAG; s->n = n;
    famg_status st = dev_alloc(&s->d, n);
    if (st != FAMG_OK) { smoother_release(s); return st; }
    cudaError_t e = cudaMemcpyAsync(s->d, d, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { smoother_release(s); FAMG_FAIL(FAMG_ERR_CUDA, "diag upload failed: %s", cudaGetErrorString(e)); }
    *out = s;
    return FAMG_OK;
}

famg_status famg_smoother_cholesky(const famg_csr *a, famg_smoother **out) {
    if (!a || !out) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    *out = nullptr;
    if (a->nrows != a->ncols) FAMG_FAIL(FAMG_ERR_INVALID, "Cholesky solve needs a square matrix");
    CUDA_TRY(cudaSetDevice(a->ctx->device));
    famg_smoother *s = new famg_smoother();
    s->ctx = a->ctx; s->kind = SM_DENSE_INV; s->n = a->nrows;
    famg_status st = dense_inverse_from_csr(a, &s->inv);
    if (st != FAMG_OK) { smoother_release(s); return st; }
    *out = s;
    return FAMG_OK;
}

// ---------------------------------------------------------------- block smoother build on the device
// BlockSmoother::new / diagonally_compensate (block_smoothers.rs:88-123, 293-324) for scalar operators.  One thread per
// aggregate: compensated principal block, Cholesky, explicit inverse -- the statements of the host routine below in the same
// order (products and sums are rounded separately on both sides: -fmad=false here, no FMA contraction there), so both builds
// produce the same bits.  The per-aggregate work arrays are interleaved across aggregates (element e of aggregate g at
// e * n_aggs + g): the threads of a warp walk the same element of neighbouring aggregates, coalesced.
constexpr int BS_DEV_MAX_AGG = 64;  // larger aggregates: host build (one thread per block would crawl)

// node -> (aggregate, position inside it); flags: 1 = node out of range or in two aggregates, 2 = nodes not ascending
__global__ void bs_node_map_kernel(const int *__restrict__ agg_ptr, const int *__restrict__ agg_nodes, int n_aggs, int n,
                                   int *__restrict__ node_agg, int *__restrict__ node_local, int *__restrict__ flags) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_aggs) return;
    const int u0 = agg_ptr[g], u1 = agg_ptr[g + 1];
    for (int u = u0; u < u1; ++u) {
        const int node = agg_nodes[u];
        if (node < 0 || node >= n) { atomicOr(flags, 1); continue; }
        if (u > u0 && agg_nodes[u - 1] >= node) atomicOr(flags, 2);
        if (atomicExch(&node_agg[node], g) != -1) atomicOr(flags, 1);
        node_local[node] = u - u0;
    }
}

// row i of the block-diagonal M^-1 has one entry per node of i's aggregate
__global__ void bs_rowlen_kernel(const int *__restrict__ agg_ptr, const int *__restrict__ node_agg, int n, int *__restrict__ len) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int g = node_agg[i];
    len[i] = g >= 0 ? agg_ptr[g + 1] - agg_ptr[g] : 0;
}

__global__ void __launch_bounds__(128) bs_build_kernel(const int *__restrict__ rp, const int *__restrict__ col, const double *__restrict__ val,
                                                       const int *__restrict__ agg_ptr, const int *__restrict__ agg_nodes,
                                                       const int *__restrict__ node_agg, const int *__restrict__ node_local,
                                                       const double *__restrict__ diag, int n_aggs, double *__restrict__ wb,
                                                       double *__restrict__ wl, const int *__restrict__ m_rp, int *__restrict__ m_col,
                                                       double *__restrict__ m_val, int *__restrict__ bad_agg) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_aggs) return;
    const size_t S = (size_t)n_aggs;
    const int u0 = agg_ptr[g], na = agg_ptr[g + 1] - u0;
    const int *nodes = agg_nodes + u0;
#define BS_B(r, c) wb[((size_t)(r) + (size_t)(c) * na) * S + g]
#define BS_L(r, c) wl[((size_t)(r) + (size_t)(c) * na) * S + g]
    for (int e = 0; e < na * na; ++e) { wb[(size_t)e * S + g] = 0.0; wl[(size_t)e * S + g] = 0.0; }
    for (int li = 0; li < na; ++li) {
        const int i = nodes[li];
        // entries in row order; the diagonal and the lumped off-aggregate couplings are summed in that order
        bool diag_seen = false;
        double dsum = 0.0;
        const double di = diag[i];
        for (int q = rp[i]; q < rp[i + 1]; ++q) {
            const int j = col[q];
            if (node_agg[j] == g) {
                const int lj = node_local[j];
                if (lj == li) { dsum = diag_seen ? dsum + val[q] : val[q]; diag_seen = true; }
                else BS_B(li, lj) = val[q];
            } else {
                const double comp = 0.5 * sqrt(di / diag[j]) * fabs(val[q]);
                dsum = diag_seen ? dsum + comp : comp; diag_seen = true;
            }
        }
        BS_B(li, li) = dsum;
    }
    // the reference factorises the upper side of the transpose == lower triangle of the block
    for (int c = 0; c < na; ++c)
        for (int r = 0; r < c; ++r) BS_B(r, c) = BS_B(c, r);
    for (int j = 0; j < na; ++j) {  // Cholesky, column by column
        double d = BS_B(j, j);
        for (int k = 0; k < j; ++k) d -= BS_L(j, k) * BS_L(j, k);
        if (!(d > 0.0)) { atomicMin(bad_agg, g); return; }
        d = sqrt(d);
        BS_L(j, j) = d;
        for (int i = j + 1; i < na; ++i) {
            double t = BS_B(i, j);
            for (int k = 0; k < j; ++k) t -= BS_L(i, k) * BS_L(j, k);
            BS_L(i, j) = t / d;
        }
    }
    for (int c = 0; c < na; ++c) {  // column c of the inverse: L L^T x = e_c
        for (int i = 0; i < na; ++i) BS_B(i, c) = i == c ? 1.0 : 0.0;
        for (int i = 0; i < na; ++i) {
            double t = BS_B(i, c);
            for (int j = 0; j < i; ++j) t -= BS_L(i, j) * BS_B(j, c);
            BS_B(i, c) = t / BS_L(i, i);
        }
        for (int i = na - 1; i >= 0; --i) {
            double t = BS_B(i, c);
            for (int j = i + 1; j < na; ++j) t -= BS_L(j, i) * BS_B(j, c);
            BS_B(i, c) = t / BS_L(i, i);
        }
    }
    for (int li = 0; li < na; ++li) {
        const int base = m_rp[nodes[li]];
        for (int lj = 0; lj < na; ++lj) { m_col[base + lj] = nodes[lj]; m_val[base + lj] = BS_B(li, lj); }
    }
#undef BS_B
#undef BS_L
}

// returns FAMG_ERR_UNSUPPORTED (without setting an error) when the partition is not one for the device build
static famg_status block_smoother_build_dev(const famg_csr *a, int64_t n_aggs, const uint64_t *agg_ptr, const uint64_t *agg_nodes,
                                            famg_smoother **out) {
    famg_ctx *ctx = a->ctx;
    const int64_t n = a->nrows;
    int64_t na_max = 0, total = 0;
    for (int64_t g = 0; g < n_aggs; ++g) {
        if (agg_ptr[g + 1] < agg_ptr[g]) FAMG_FAIL(FAMG_ERR_INVALID, "invalid partition");
        const int64_t na = (int64_t)(agg_ptr[g + 1] - agg_ptr[g]);
        na_max = std::max(na_max, na); total += na * na;
    }
    // interleaved work arrays hold na_max^2 elements for every aggregate: only for partitions of similar, small aggregates
    if (n == 0 || n_aggs == 0 || na_max > BS_DEV_MAX_AGG || total > 0x7fffffff || na_max * na_max * n_aggs > 2 * total + 4096)
        return FAMG_ERR_UNSUPPORTED;
    std::vector<int> h_ptr((size_t)n_aggs + 1), h_nodes((size_t)n);
    for (int64_t g = 0; g <= n_aggs; ++g) h_ptr[(size_t)g] = (int)agg_ptr[g];
    for (int64_t u = 0; u < n; ++u) {
        if (agg_nodes[u] >= (uint64_t)n) FAMG_FAIL(FAMG_ERR_INVALID, "invalid partition");
        h_nodes[(size_t)u] = (int)agg_nodes[u];
    }
    const size_t work = (size_t)(na_max * na_max) * (size_t)n_aggs;
    int *d_ptr = nullptr, *d_nodes = nullptr, *d_map = nullptr, *d_flags = nullptr;
    double *d_diag = nullptr, *wb = nullptr, *wl = nullptr;
    famg_csr *minv = nullptr;
    auto cleanup = [&]() {
        pool_free(ctx, d_ptr, 0); pool_free(ctx, d_nodes, 0); pool_free(ctx, d_map, 0); pool_free(ctx, d_flags, 0);
        pool_free(ctx, d_diag, 0); pool_free(ctx, wb, 0); pool_free(ctx, wl, 0);
    };
#define BS_TRY(expr) do { famg_status s__ = (expr); if (s__ != FAMG_OK) { cleanup(); if (minv) csr_release(minv); return s__; } } while (0)
#define BS_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { set_error("block smoother build: %s", cudaGetErrorString(e__)); BS_TRY(FAMG_ERR_CUDA); } } while (0)
    BS_TRY(pool_alloc(ctx, sizeof(int) * ((size_t)n_aggs + 1), (void **)&d_ptr));
    BS_TRY(pool_alloc(ctx, sizeof(int) * (size_t)n, (void **)&d_nodes));
    BS_TRY(pool_alloc(ctx, sizeof(int) * (size_t)3 * (size_t)n, (void **)&d_map));  // node_agg | node_local | row lengths
    BS_TRY(pool_alloc(ctx, 256, (void **)&d_flags));                               // [0] partition flags, [1] missing diagonals, [2] first non-SPD aggregate
    BS_TRY(pool_alloc(ctx, sizeof(double) * (size_t)n, (void **)&d_diag));
    BS_TRY(pool_alloc(ctx, sizeof(double) * work, (void **)&wb));
    BS_TRY(pool_alloc(ctx, sizeof(double) * work, (void **)&wl));
    int *node_agg = d_map, *node_local = d_map + n, *row_len = d_map + 2 * n;
    const int h_init[3] = {0, 0, 0x7fffffff};
    BS_CUDA(cudaMemcpyAsync(d_ptr, h_ptr.data(), sizeof(int) * ((size_t)n_aggs + 1), cudaMemcpyHostToDevice, ctx->stream));
    BS_CUDA(cudaMemcpyAsync(d_nodes, h_nodes.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    BS_CUDA(cudaMemcpyAsync(d_flags, h_init, sizeof(h_init), cudaMemcpyHostToDevice, ctx->stream));
    BS_CUDA(cudaMemsetAsync(node_agg, 0xff, sizeof(int) * (size_t)n, ctx->stream));
    const unsigned grid_n = (unsigned)ceil_div(n, 256), grid_g = (unsigned)ceil_div(n_aggs, 128);
    bs_node_map_kernel<<<grid_g, 128, 0, ctx->stream>>>(d_ptr, d_nodes, (int)n_aggs, (int)n, node_agg, node_local, d_flags);
    bs_rowlen_kernel<<<grid_n, 256, 0, ctx->stream>>>(d_ptr, node_agg, (int)n, row_len);
    diag_extract_linear_kernel<<<grid_n, 256, 0, ctx->stream>>>(a->row_ptr, a->col, a->val, (int)n, d_diag, d_flags + 1);
    count_launch(ctx, 3);
    BS_CUDA(cudaGetLastError());
    int h_flags[3] = {0, 0, 0};
    BS_CUDA(cudaMemcpyAsync(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, ctx->stream));
    BS_CUDA(cudaStreamSynchronize(ctx->stream));  // also: h_ptr / h_nodes / h_init have been read
    if (h_flags[0] & 1) { set_error("invalid partition"); BS_TRY(FAMG_ERR_INVALID); }
    if (h_flags[0] & 2) { set_error("aggregate nodes must ascend"); BS_TRY(FAMG_ERR_INVALID); }
    if (h_flags[1]) { set_error("%d rows have no diagonal entry", h_flags[1]); BS_TRY(FAMG_ERR_NUMERIC); }
    BS_TRY(csr_alloc(ctx, n, n, total, &minv));
    BS_TRY(exclusive_scan_i32(ctx, row_len, minv->row_ptr, n));
    bs_build_kernel<<<grid_g, 128, 0, ctx->stream>>>(a->row_ptr, a->col, a->val, d_ptr, d_nodes, node_agg, node_local, d_diag, (int)n_aggs, wb, wl,
                                                     minv->row_ptr, minv->col, minv->val, d_flags + 2);
    count_launch(ctx);
    BS_CUDA(cudaGetLastError());
    BS_CUDA(cudaMemcpyAsync(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, ctx->stream));
    BS_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h_flags[2] != 0x7fffffff) { set_error("aggregate %d block is not positive definite", h_flags[2]); BS_TRY(FAMG_ERR_NUMERIC); }
    BS_TRY(csr_finalize_plan(minv));
    cleanup();
#undef BS_TRY
#undef BS_CUDA
    famg_smoother *s = new famg_smoother();
    s->ctx = ctx; s->kind = SM_SPARSE_INV; s->n = n; s->minv = minv;
    *out = s;
    return FAMG_OK;
}

// Dense SPD inverse of a small block on the host (Cholesky, then solve for the identity).
static bool host_spd_inverse(int n, std::vector<double> &m /* n x n col-major, in: A, out: A^-1 */) {
    std::vector<double> l((size_t)n * n, 0.0);
    for (int j = 0; j < n; ++j) {
        double d = m[(size_t)j + (size_t)j * n];
        for (int k = 0; k < j; ++k) d -= l[(size_t)j + (size_t)k * n] * l[(size_t)j + (size_t)k * n];
        if (!(d > 0.0)) return false;
        d = sqrt(d);
        l[(size_t)j + (size_t)j * n] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = m[(size_t)i + (size_t)j * n];
            for (int k = 0; k < j; ++k) s -= l[(size_t)i + (size_t)k * n] * l[(size_t)j + (size_t)k * n];
            l[(size_t)i + (size_t)j * n] = s / d;
        }
    }
    for (int c = 0; c < n; ++c) {
        double *x = &m[(size_t)c * n];
        for (int i = 0; i < n; ++i) x[i] = i == c ? 1.0 : 0.0;
        for (int i = 0; i < n; ++i) {
            double s = x[i];
            for (int j = 0; j < i; ++j) s -= l[(size_t)i + (size_t)j * n] * x[j];
            x[i] = s / l[(size_t)i + (size_t)i * n];
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = x[i];
            for (int j = i + 1; j < n; ++j) s -= l[(size_t)j + (size_t)i * n] * x[j];
            x[i] = s / l[(size_t)i + (size_t)i * n];
        }
    }
    return true;
}

famg_status famg_smoother_block(const famg_csr *a, int64_t n_aggs, const uint64_t *agg_ptr, const uint64_t *agg_nodes,
                                famg_smoother **out) {
    if (!a || !out || !agg_ptr || !agg_nodes || n_aggs < 0) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    *out = nullptr;
    if (a->nrows != a->ncols) FAMG_FAIL(FAMG_ERR_INVALID, "block smoother needs a square matrix");
    if ((int64_t)agg_ptr[n_aggs] != a->nrows) FAMG_FAIL(FAMG_ERR_INVALID, "partition does not cover the matrix");  // block_smoothers.rs:91
    CUDA_TRY(cudaSetDevice(a->ctx->device));
    if (!a->ctx->block_build_host) {  // aggregates of at most BS_DEV_MAX_AGG nodes: built on the device, same bits
        const famg_status ds = block_smoother_build_dev(a, n_aggs, agg_ptr, agg_nodes, out);
        if (ds != FAMG_ERR_UNSUPPORTED) return ds;
    }
    HostCsr h;
    FAMG_TRY(csr_to_host(a, &h));
    const int64_t n = a->nrows;
    std::vector<double> diag((size_t)n, 0.0);
    for (int64_t i = 0; i < n; ++i) {
        bool found = false;
        for (int q = h.row_ptr[(size_t)i]; q < h.row_ptr[(size_t)i + 1]; ++q)
            if (h.col[(size_t)q] == i) { diag[(size_t)i] = h.val[(size_t)q]; found = true; break; }
        if (!found) FAMG_FAIL(FAMG_ERR_NUMERIC, "row %lld has no diagonal entry", (long long)i);
    }
    // per aggregate: compensated principal block -> inverse; scatter as rows of block-diagonal M^-1
    std::vector<int64_t> node_agg((size_t)n, -1), node_local((size_t)n, 0);
    for (int64_t g = 0; g < n_aggs; ++g)
        for (uint64_t u = agg_ptr[g]; u < agg_ptr[g + 1]; ++u) {
            uint64_t node = agg_nodes[u];
            if (node >= (uint64_t)n || node_agg[(size_t)node] != -1) FAMG_FAIL(FAMG_ERR_INVALID, "invalid partition");
            if (u > agg_ptr[g] && agg_nodes[u - 1] >= node) FAMG_FAIL(FAMG_ERR_INVALID, "aggregate nodes must ascend");
            node_agg[(size_t)node] = g; node_local[(size_t)node] = (int64_t)(u - agg_ptr[g]);
        }
    std::vector<int> rp((size_t)n + 1, 0);
    for (int64_t i = 0; i < n; ++i) {
        int64_t g = node_agg[(size_t)i];
        rp[(size_t)i + 1] = (int)(agg_ptr[g + 1] - agg_ptr[g]);
    }
    for (int64_t i = 0; i < n; ++i) rp[(size_t)i + 1] += rp[(size_t)i];
    std::vector<int> ci((size_t)rp[(size_t)n]);
    std::vector<double> cv((size_t)rp[(size_t)n]);
    std::vector<double> blk;
    for (int64_t g = 0; g < n_aggs; ++g) {
        const int na = (int)(agg_ptr[g + 1] - agg_ptr[g]);
        const uint64_t *nodes = agg_nodes + agg_ptr[g];
        blk.assign((size_t)na * na, 0.0);
        for (int li = 0; li < na; ++li) {
            const int64_t i = (int64_t)nodes[li];
            // triplets in row order; duplicates on (li,li) summed in order (try_new_from_triplets)
            bool diag_seen = false; double dsum = 0.0;
            for (int q = h.row_ptr[(size_t)i]; q < h.row_ptr[(size_t)i + 1]; ++q) {
                const int64_t j = h.col[(size_t)q];
                if (node_agg[(size_t)j] == g) {
                    const int lj = (int)node_local[(size_t)j];
                    if (lj == li) { dsum = diag_seen ? dsum + h.val[(size_t)q] : h.val[(size_t)q]; diag_seen = true; }
                    else blk[(size_t)li + (size_t)lj * na] = h.val[(size_t)q];
                } else {
                    const double comp = 0.5 * sqrt(diag[(size_t)i] / diag[(size_t)j]) * fabs(h.val[(size_t)q]);
                    dsum = diag_seen ? dsum + comp : comp; diag_seen = true;
                }
            }
            blk[(size_t)li + (size_t)li * na] = dsum;
        }
        // the reference factorises the upper side of the transpose == lower triangle of the block
        for (int c = 0; c < na; ++c)
            for (int r = 0; r < c; ++r) blk[(size_t)r + (size_t)c * na] = blk[(size_t)c + (size_t)r * na];
        if (!host_spd_inverse(na, blk)) FAMG_FAIL(FAMG_ERR_NUMERIC, "aggregate %lld block is not positive definite", (long long)g);
        for (int li = 0; li < na; ++li) {
            const int64_t i = (int64_t)nodes[li];
            for (int lj = 0; lj < na; ++lj) {
                ci[(size_t)rp[(size_t)i] + lj] = (int)nodes[lj];
                cv[(size_t)rp[(size_t)i] + lj] = blk[(size_t)li + (size_t)lj * na];
            }
        }
    }
    famg_smoother *s = new famg_smoother();
    s->ctx = a->ctx; s->kind = SM_SPARSE_INV; s->n = n;
    famg_status st = csr_from_host_i32(a->ctx, n, n, rp.data(), ci.data(), cv.data(), &s->minv);
    if (st != FAMG_OK) { smoother_release(s); return st; }
    *out = s;
    return FAMG_OK;
}

// BlockSmoother::new for vdim > 1 (block_smoothers.rs:88-123 with diagonally_compensate_vector, :326-400):
// the partition is over *nodes* of vdim dofs each.  Per aggregate: couplings to nodes of the aggregate are
// kept; every coupling block A_IJ to a node J outside is lumped into node I's diagonal block as
// 0.5 * U S U^T with -A_IJ = U S V^T (the symmetric polar factor; independent of the SVD's sign choices).
// The reference walks a HashSet of the blocks to lump (arbitrary order); here they are added in ascending
// (I, J) order.  The block is solved exactly (Cholesky) and materialised as rows of the block-diagonal M^-1.
famg_status famg_smoother_block_vector(const famg_csr *a, int64_t vdim, int64_t n_aggs, const uint64_t *agg_ptr,
                                       const uint64_t *agg_nodes, famg_smoother **out) {
    if (!a || !out || !agg_ptr || !agg_nodes || n_aggs < 0 || vdim < 1 || vdim > 64) FAMG_FAIL(FAMG_ERR_INVALID, "bad argument");
    if (vdim == 1) return famg_smoother_block(a, n_aggs, agg_ptr, agg_nodes, out);  // block_smoothers.rs:98-102
    *out = nullptr;
    if (a->nrows != a->ncols || a->nrows % vdim != 0) FAMG_FAIL(FAMG_ERR_INVALID, "block smoother needs a square matrix of vdim-sized nodes");
    const int64_t n = a->nrows, nnodes = n / vdim;
    if ((int64_t)agg_ptr[n_aggs] != nnodes) FAMG_FAIL(FAMG_ERR_INVALID, "partition does not cover the matrix");  // block_smoothers.rs:91
    CUDA_TRY(cudaSetDevice(a->ctx->device));
    HostCsr h;
    FAMG_TRY(csr_to_host(a, &h));
    std::vector<int64_t> node_agg((size_t)nnodes, -1), node_local((size_t)nnodes, 0);
    for (int64_t g = 0; g < n_aggs; ++g)
        for (uint64_t u = agg_ptr[g]; u < agg_ptr[g + 1]; ++u) {
            const uint64_t node = agg_nodes[u];
            if (node >= (uint64_t)nnodes || node_agg[(size_t)node] != -1) FAMG_FAIL(FAMG_ERR_INVALID, "invalid partition");
            if (u > agg_ptr[g] && agg_nodes[u - 1] >= node) FAMG_FAIL(FAMG_ERR_INVALID, "aggregate nodes must ascend");
            node_agg[(size_t)node] = g; node_local[(size_t)node] = (int64_t)(u - agg_ptr[g]);
        }
    std::vector<int> rp((size_t)n + 1, 0);
    for (int64_t i = 0; i < n; ++i) {
        const int64_t g = node_agg[(size_t)(i / vdim)];
        rp[(size_t)i + 1] = rp[(size_t)i] + (int)((agg_ptr[g + 1] - agg_ptr[g]) * (uint64_t)vdim);
    }
    std::vector<int> ci((size_t)rp[(size_t)n]);
    std::vector<double> cv((size_t)rp[(size_t)n]);
    const int vd = (int)vdim;
    int failed = 0;
#pragma omp parallel
    {
        std::vector<double> blk, diag, u, sv, vv, usut((size_t)vd * vd);
        std::vector<int64_t> outside;
#pragma omp for schedule(dynamic, 16)
        for (int64_t g = 0; g < n_aggs; ++g) {
            const int na = (int)(agg_ptr[g + 1] - agg_ptr[g]);
            const int bd = na * vd;
            const uint64_t *nodes = agg_nodes + agg_ptr[g];
            blk.assign((size_t)bd * bd, 0.0);
            diag.assign((size_t)na * vd * vd, 0.0);  // per node, row-major vd x vd
            for (int li = 0; li < na; ++li) {
                const int64_t bi = (int64_t)nodes[li];
                outside.clear();
                for (int oi = 0; oi < vd; ++oi) {
                    const int64_t i = bi * vd + oi;
                    for (int q = h.row_ptr[(size_t)i]; q < h.row_ptr[(size_t)i + 1]; ++q) {
                        const int64_t j = h.col[(size_t)q], bj = j / vd;
                        const int oj = (int)(j % vd);
                        if (bj == bi) diag[((size_t)li * vd + oi) * vd + oj] = h.val[(size_t)q];
                        else if (node_agg[(size_t)bj] == g) blk[(size_t)(li * vd + oi) + (size_t)(node_local[(size_t)bj] * vd + oj) * bd] = h.val[(size_t)q];
                        else outside.push_back(bj);
                    }
                }
                std::sort(outside.begin(), outside.end());
                outside.erase(std::unique(outside.begin(), outside.end()), outside.end());
                for (int64_t bj : outside) {
                    u.assign((size_t)vd * vd, 0.0);  // -A_IJ, column-major
                    for (int oi = 0; oi < vd; ++oi) {
                        const int64_t i = bi * vd + oi;
                        for (int q = h.row_ptr[(size_t)i]; q < h.row_ptr[(size_t)i + 1]; ++q) {
                            const int64_t j = h.col[(size_t)q];
                            if (j / vd == bj) u[(size_t)oi + (size_t)(j % vd) * vd] -= h.val[(size_t)q];
                        }
                    }
                    host_thin_svd(vd, vd, u, sv, vv);
                    for (int r = 0; r < vd; ++r)
                        for (int c = 0; c < vd; ++c) {
                            double t = 0.0;
                            for (int e = 0; e < vd; ++e) t += u[(size_t)r + (size_t)e * vd] * (sv[(size_t)e] * u[(size_t)c + (size_t)e * vd]);
                            usut[(size_t)r * vd + c] = t;
                        }
                    for (int e = 0; e < vd * vd; ++e) diag[(size_t)li * vd * vd + e] += 0.5 * usut[(size_t)e];
                }
                for (int oi = 0; oi < vd; ++oi)
                    for (int oj = 0; oj < vd; ++oj) blk[(size_t)(li * vd + oi) + (size_t)(li * vd + oj) * bd] = diag[((size_t)li * vd + oi) * vd + oj];
            }
            // the reference factorises the upper side of the transpose == lower triangle of the block
            for (int c = 0; c < bd; ++c)
                for (int r = 0; r < c; ++r) blk[(size_t)r + (size_t)c * bd] = blk[(size_t)c + (size_t)r * bd];
            if (!host_spd_inverse(bd, blk)) {
#pragma omp atomic write
                failed = 1;
                continue;
            }
            for (int li = 0; li < na; ++li)
                for (int oi = 0; oi < vd; ++oi) {
                    const int64_t i = (int64_t)nodes[li] * vd + oi;
                    for (int lj = 0; lj < na; ++lj)
                        for (int oj = 0; oj < vd; ++oj) {
                            ci[(size_t)rp[(size_t)i] + lj * vd + oj] = (int)((int64_t)nodes[lj] * vd + oj);
                            cv[(size_t)rp[(size_t)i] + lj * vd + oj] = blk[(size_t)(li * vd + oi) + (size_t)(lj * vd + oj) * bd];
                        }
                }
        }
    }
    if (failed) FAMG_FAIL(FAMG_ERR_NUMERIC, "an aggregate block is not positive definite");
    famg_smoother *s = new famg_smoother();
    s->ctx = a->ctx; s->kind = SM_SPARSE_INV; s->n = n;
    famg_status st = csr_from_host_i32(a->ctx, n, n, rp.data(), ci.data(), cv.data(), &s->minv);
    if (st != FAMG_OK) { smoother_release(s); return st; }
    *out = s;
    return FAMG_OK;
}

famg_status famg_smoother_retain(famg_smoother *s) {
    if (!s) FAMG_FAIL(FAMG_ERR_INVALID, "null smoother");
    s->refs.fetch_add(1);
    return FAMG_OK;
}
famg_status famg_smoother_destroy(famg_smoother *s) {
    if (s) { cudaSetDevice(s->ctx->device); smoother_release(s); }
    return FAMG_OK;
}
famg_status famg_smoother_dim(const famg_smoother *s, int64_t *n) {
    if (!s || !n) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    *n = s->n;
    return FAMG_OK;
}
famg_status famg_smoother_diag_download(const famg_smoother *s, double *d) {
    if (!s || !d) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    if (s->kind != SM_DIAG) FAMG_FAIL(FAMG_ERR_INVALID, "not a diagonal smoother");
    CUDA_TRY(cudaSetDevice(s->ctx->device));
    CUDA_TRY(cudaStreamSynchronize(s->ctx->stream));
    CUDA_TRY(cudaMemcpy(d, s->d, sizeof(double) * s->n, cudaMemcpyDeviceToHost));
    return FAMG_OK;
}

famg_status famg_smoother_apply_dev(const famg_smoother *s, famg_vec *out, const famg_vec *rhs) {
    if (!s || !out || !rhs) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    if (rhs->nrows != s->n || out->nrows != s->n || out->ncols != rhs->ncols) FAMG_FAIL(FAMG_ERR_INVALID, "smoother apply shape mismatch");
    CUDA_TRY(cudaSetDevice(s->ctx->device));
    return smoother_apply_dev(s, rhs->p, rhs->ld, out->p, out->ld, (int)rhs->ncols);
}

famg_status famg_smoother_apply(const famg_smoother *s, double *out, int64_t ld_out, const double *rhs, int64_t ld_rhs, int64_t k) {
    if (!s || !out || !rhs || k < 0) FAMG_FAIL(FAMG_ERR_INVALID, "bad argument");
    if (k == 0) return FAMG_OK;
    famg_vec *x = nullptr, *y = nullptr;
    FAMG_TRY(famg_vec_create(s->ctx, s->n, k, &x));
    famg_status st = famg_vec_create(s->ctx, s->n, k, &y);
    if (st == FAMG_OK) st = famg_vec_upload(x, rhs, ld_rhs);
    if (st == FAMG_OK) st = famg_smoother_apply_dev(s, y, x);
    if (st == FAMG_OK) st = famg_vec_download(y, out, ld_out);
    famg_vec_destroy(x); famg_vec_destroy(y);
    return st;
}

// smooth(): iters x { x += M^-1 (b - A x) }.  Diag: one fused kernel per sweep (ping-pong buffer).
famg_status famg_smooth_dev(const famg_csr *a, const famg_smoother *s, famg_vec *x, const famg_vec *b, int iters) {
    if (!a || !s || !x || !b || iters < 0) FAMG_FAIL(FAMG_ERR_INVALID, "bad argument");
    if (a->nrows != a->ncols || s->n != a->nrows || x->nrows != a->nrows || b->nrows != a->nrows || b->ncols != x->ncols)
        FAMG_FAIL(FAMG_ERR_INVALID, "smooth shape mismatch");
    famg_ctx *ctx = a->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (iters == 0) return FAMG_OK;
    famg_vec *t = nullptr, *t2 = nullptr;
    FAMG_TRY(famg_vec_create(ctx, x->nrows, x->ncols, &t));
    famg_status st = FAMG_OK;
    const int k = (int)x->ncols;
    double *cur = x->p; int64_t ldc = x->ld;
    double *oth = t->p; int64_t ldo = t->ld;
    if (s->kind != SM_DIAG) st = famg_vec_create(ctx, x->nrows, x->ncols, &t2);
    for (int it = 0; it < iters && st == FAMG_OK; ++it) {
        if (s->kind == SM_DIAG) {
            SpmvArgs g; g.a = a; g.epi = EPI_SMOOTH; g.x = cur; g.ldx = ldc; g.y = oth; g.ldy = ldo; g.b = b->p; g.ldb = b->ld;
            g.d = s->d; g.k = k;
            st = spmv_launch(g);
            std::swap(cur, oth); std::swap(ldc, ldo);
        } else {
            // r = b - A x ; z = M^-1 r ; x += z
            SpmvArgs g; g.a = a; g.epi = EPI_RESID; g.x = cur; g.ldx = ldc; g.y = oth; g.ldy = ldo; g.b = b->p; g.ldb = b->ld; g.k = k;
            st = spmv_launch(g);
            if (st == FAMG_OK) st = smoother_apply_dev(s, oth, ldo, t2->p, t2->ld, k);
            for (int c = 0; c < k && st == FAMG_OK; ++c) st = vec_add_inplace(ctx, cur + c * ldc, t2->p + c * t2->ld, x->nrows);
        }
    }
    if (st == FAMG_OK && cur != x->p) st = vec_copy(ctx, x->p, x->ld, cur, ldc, x->nrows, k);
    famg_vec_destroy(t);  // returns the block to the context pool (stream-ordered reuse)
    famg_vec_destroy(t2);
    return st;
}

famg_status famg_stationary_iteration_dev(const famg_csr *a, const famg_smoother *diag, int iters, famg_vec *io) {
    if (!a || !diag || !io) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    if (diag->kind != SM_DIAG) FAMG_FAIL(FAMG_ERR_UNSUPPORTED, "stationary iteration supports Diag preconditioners");
    if (a->nrows != a->ncols || diag->n != a->nrows || io->nrows != a->nrows) FAMG_FAIL(FAMG_ERR_INVALID, "shape mismatch");
    famg_ctx *ctx = a->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    famg_vec *t = nullptr;
    FAMG_TRY(famg_vec_create(ctx, io->nrows, io->ncols, &t));
    const int k = (int)io->ncols;
    // x = M rhs
    famg_status st = vec_scale_rows(ctx, diag->d, io->p, io->ld, t->p, t->ld, io->nrows, k);
    double *cur = t->p; int64_t ldc = t->ld;
    double *oth = io->p; int64_t ldo = io->ld;
    for (int it = 1; it < iters && st == FAMG_OK; ++it) {
        SpmvArgs g; g.a = a; g.epi = EPI_SI; g.x = cur; g.ldx = ldc; g.y = oth; g.ldy = ldo; g.d = diag->d; g.k = k;
        st = spmv_launch(g);
        std::swap(cur, oth); std::swap(ldc, ldo);
    }
    if (st == FAMG_OK && cur != io->p) st = vec_copy(ctx, io->p, io->ld, cur, ldc, io->nrows, k);
    famg_vec_destroy(t);
    return st;
}

}  // extern "C"
