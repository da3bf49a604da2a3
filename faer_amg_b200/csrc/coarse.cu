// coarse.cu -- coarsest-level exact solve.
// Replaces SparseCholeskySolve::{new, apply} (coarse_solvers.rs:173-206) for the coarsest level
// (multigrid.rs:105-109).  n_c <= coarsest_dim (1000 by default, hierarchy.rs:31), so the factor
// is computed once on the host and the solve is realised as x = A_c^{-1} f, a dense f64 GEMV that
// streams the (L2-resident) inverse once: two dependent triangular solves would be pure latency on
// a GPU.
#include <cmath>

#include "common.cuh"

namespace famg {

constexpr int64_t DENSE_MAX = 8192;

famg_status dense_inverse_from_csr(const famg_csr *a, double **d_inv) {
    *d_inv = nullptr;
    const int64_t n = a->nrows;
    if (n > DENSE_MAX) FAMG_FAIL(FAMG_ERR_UNSUPPORTED, "coarse solve of dimension %lld exceeds the dense limit %lld", (long long)n, (long long)DENSE_MAX);
    HostCsr h;
    FAMG_TRY(csr_to_host(a, &h));
    std::vector<double> l((size_t)n * n, 0.0);  // column-major; lower triangle of A, then L
    for (int64_t i = 0; i < n; ++i)
        for (int q = h.row_ptr[(size_t)i]; q < h.row_ptr[(size_t)i + 1]; ++q)
            if (h.col[(size_t)q] <= i) l[(size_t)i + (size_t)h.col[(size_t)q] * n] = h.val[(size_t)q];
    // right-looking Cholesky (column-oriented, cache friendly)
    for (int64_t j = 0; j < n; ++j) {
        double d = l[(size_t)j + (size_t)j * n];
        if (!(d > 0.0)) FAMG_FAIL(FAMG_ERR_NUMERIC, "coarse matrix is not positive definite (pivot %lld = %g)", (long long)j, d);
        d = sqrt(d);
        double *cj = &l[(size_t)j * n];
        cj[j] = d;
        for (int64_t i = j + 1; i < n; ++i) cj[i] /= d;
        // trailing update: columns are independent (same arithmetic per entry in any thread count)
#pragma omp parallel for schedule(static) if (n - j > 128)
        for (int64_t k = j + 1; k < n; ++k) {
            const double f = cj[k];
            if (f == 0.0) continue;
            double *ck = &l[(size_t)k * n];
            for (int64_t i = k; i < n; ++i) ck[i] -= cj[i] * f;
        }
    }
    // inverse: solve L L^T X = I column by column (forward solve starts at the unit row)
    std::vector<double> inv((size_t)n * n, 0.0);
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t c = 0; c < n; ++c) {
        double *x = &inv[(size_t)c * n];
        x[c] = 1.0;
        for (int64_t j = c; j < n; ++j) {  // column-oriented forward substitution
            const double xj = x[j] / l[(size_t)j + (size_t)j * n];
            x[j] = xj;
            if (xj == 0.0) continue;
            const double *cj = &l[(size_t)j * n];
            for (int64_t i = j + 1; i < n; ++i) x[i] -= cj[i] * xj;
        }
        for (int64_t j = n - 1; j >= 0; --j) {  // L^T x = y: dot with column j of L
            const double *cj = &l[(size_t)j * n];
            double s = x[j];
            for (int64_t i = j + 1; i < n; ++i) s -= cj[i] * x[i];
            x[j] = s / cj[j];
        }
    }
    // symmetrise (A^-1 is symmetric; removes rounding skew so the V-cycle stays symmetric)
    for (int64_t c = 0; c < n; ++c)
        for (int64_t r = 0; r < c; ++r) {
            const double m = 0.5 * (inv[(size_t)r + (size_t)c * n] + inv[(size_t)c + (size_t)r * n]);
            inv[(size_t)r + (size_t)c * n] = m; inv[(size_t)c + (size_t)r * n] = m;
        }
    FAMG_TRY(dev_alloc(d_inv, n * n));
    cudaError_t e = cudaMemcpy(*d_inv, inv.data(), sizeof(double) * (size_t)(n * n), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(*d_inv); *d_inv = nullptr; FAMG_FAIL(FAMG_ERR_CUDA, "inverse upload failed: %s", cudaGetErrorString(e)); }
    return FAMG_OK;
}

// y[:,c] = M x[:,c], M symmetric n x n column-major: row i of M == column i, so a warp streams
// column i with coalesced loads.  One warp per output row, fixed shuffle tree (deterministic).
__global__ void __launch_bounds__(256) dense_symv_kernel(const double *__restrict__ m, long long n, const double *__restrict__ x,
                                                         long long ldx, double *__restrict__ y, long long ldy, int k) {
    const int warp = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (warp >= n) return;
    const double *__restrict__ col = m + (long long)warp * n;
    for (int c = 0; c < k; ++c) {
        const double *__restrict__ xc = x + c * ldx;
        double s = 0.0;
        for (long long j = lane; j < n; j += 32) s += col[j] * xc[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) y[c * ldy + warp] = s;
    }
}

famg_status dense_gemv(famg_ctx *ctx, const double *inv, int64_t n, const double *x, int64_t ldx, double *y, int64_t ldy, int k,
                       cudaStream_t st) {
    if (n == 0 || k == 0) return FAMG_OK;
    if (x == y) FAMG_FAIL(FAMG_ERR_INVALID, "dense solve cannot run in place");
    const unsigned grid = (unsigned)ceil_div(n * 32, 256);
    dense_symv_kernel<<<grid, 256, 0, st ? st : ctx->stream>>>(inv, n, x, ldx, y, ldy, k);
    count_launch(ctx);
    KERNEL_CHECK();
    return FAMG_OK;
}

}  // namespace famg
