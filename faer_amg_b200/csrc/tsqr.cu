// tsqr.cu -- the near-null search loop kept resident in HBM.
//
// Reference call sites: `smooth_vector` (src/adaptivity.rs:307-390) -- one `ErrorPropogator::apply`
// (adaptivity.rs:191-198) plus one `x.qr().compute_thin_Q()` per step on an n x k block (:351-354),
// then per-column convergence factors (:365-384); and the coarse near-null QR of
// src/hierarchy.rs:228.  The reference does a Householder QR on the host; doing the same here would
// move 2*n*k*8 bytes over PCIe per step, so the block stays on the device:
//   thin Q        CholeskyQR2:  G = X^T X (deterministic two-stage reduction) -> k x k Cholesky on
//                 the host (k <= 64: microseconds) -> X <- X R^-1, twice.  Q is the same matrix the
//                 Householder QR produces (R with a positive diagonal, as famg_thin_q) up to
//                 O(eps * cond(X)^2) in the first pass, O(eps) after the second.
//   E apply       Diag preconditioner: one fused SpMM launch, out = x - d .* (A x) (EPI_EPROP);
//                 any other smoother: SpMM, smoother apply, axpby.
//   column dots   w^T A w, (Ev)^T A (Ev): deterministic two-stage reductions.
// The dense kernels use explicit fma() -- there is no oracle bit-parity to keep for a QR that is
// algorithmically different from the reference's; the SpMM keeps the library's -fmad=false order.
//
// Roofline: all kernels stream the n x k block (8nk bytes per pass).  Per step: E apply reads A and
// x, writes x' (12 nnz + 8n + 16nk B); each CholeskyQR pass reads x (Gram) and reads + writes it
// (R^-1): 24nk B per pass, 48nk per QR.  The Gram and R^-1 kernels do n k^2 (resp. n k^2 / 2)
// FP64 FMAs out of shared memory with 4x4 (resp. 1x16) register blocking so that at k = 64 they stay
// at the FP64 pipe's rate rather than the shared-memory port's.
#include <cmath>

#include "common.cuh"

namespace famg {

constexpr int QR_MAXK = 64;
constexpr int QR_TILE = 64;      // rows per tile
constexpr int QR_THREADS = 256;
constexpr int QR_LDT = QR_MAXK + 1;  // padded tile row: column reads by consecutive rows are conflict-free

// stage rows [r0, r0 + QR_TILE) of the first k columns: tile[r][c]; consecutive threads read
// consecutive rows of one column (coalesced)
__device__ __forceinline__ void qr_load_tile(double (*tile)[QR_LDT], const double *__restrict__ x, long long ld, long long n,
                                             long long r0, int k) {
    for (int t = threadIdx.x; t < QR_TILE * k; t += QR_THREADS) {
        const int r = t % QR_TILE, c = t / QR_TILE;
        tile[r][c] = r0 + r < n ? x[(long long)c * ld + r0 + r] : 0.0;
    }
}

// partial Gram matrices, one per CTA: partials[blockIdx][i + j*k].  KB = ceil(k / 16).  Thread (ti, tj)
// of a 16 x 16 grid owns the KB x KB entries {ti + 16a} x {tj + 16b}: per tile row 2 KB shared loads
// (consecutive ti -> consecutive words; tj takes 2 values per warp -> broadcast) feed KB^2 FMAs.
template <int KB>
__global__ void __launch_bounds__(QR_THREADS, 3) gram_partial_kernel(const double *__restrict__ x, long long ld, long long n, int k,
                                                                  double *__restrict__ partials) {
    __shared__ double tile[QR_TILE][QR_LDT];
    for (int t = threadIdx.x; t < QR_TILE * QR_LDT; t += QR_THREADS) (&tile[0][0])[t] = 0.0;  // columns >= k stay zero
    const int ti = threadIdx.x & 15, tj = threadIdx.x >> 4;
    double acc[KB][KB];
#pragma unroll
    for (int a = 0; a < KB; ++a)
#pragma unroll
        for (int b = 0; b < KB; ++b) acc[a][b] = 0.0;
    for (long long r0 = (long long)blockIdx.x * QR_TILE; r0 < n; r0 += (long long)gridDim.x * QR_TILE) {
        __syncthreads();
        qr_load_tile(tile, x, ld, n, r0, k);
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < QR_TILE; ++r) {
            double xi[KB], xj[KB];
#pragma unroll
            for (int a = 0; a < KB; ++a) { xi[a] = tile[r][ti + 16 * a]; xj[a] = tile[r][tj + 16 * a]; }
#pragma unroll
            for (int a = 0; a < KB; ++a)
#pragma unroll
                for (int b = 0; b < KB; ++b) acc[a][b] = fma(xi[a], xj[b], acc[a][b]);
        }
    }
#pragma unroll
    for (int a = 0; a < KB; ++a)
#pragma unroll
        for (int b = 0; b < KB; ++b) {
            const int i = ti + 16 * a, j = tj + 16 * b;
            if (i < k && j < k) partials[(long long)blockIdx.x * k * k + i + j * k] = acc[a][b];
        }
}

// out[idx] = sum over parts, fixed order (deterministic); one thread per entry
__global__ void sum_partials_kernel(const double *__restrict__ partials, int nparts, int count, double *__restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += partials[(long long)p * count + idx];
    out[idx] = s;
}

// X <- X * Rinv (upper triangular; passed ROW-major and zero-padded to QR_MAXK columns), in place.
// Thread (r = tid % 64, g = tid / 64) owns row r and the CPT = 4 KB columns [g CPT, (g + 1) CPT): per
// inner index one tile load and CPT warp-uniform (broadcast) coefficient loads feed CPT FMAs.
template <int KB>
__global__ void __launch_bounds__(QR_THREADS, 3) apply_rinv_kernel(double *__restrict__ x, long long ld, long long n, int k,
                                                                const double *__restrict__ rinv_rm) {
    constexpr int CPT = 4 * KB;
    extern __shared__ __align__(16) double qr_smem[];  // tile, then Rinv (k x QR_MAXK)
    double (*tile)[QR_LDT] = reinterpret_cast<double (*)[QR_LDT]>(qr_smem);
    double *rs = qr_smem + QR_TILE * QR_LDT;
    for (int t = threadIdx.x; t < k * QR_MAXK; t += QR_THREADS) rs[t] = rinv_rm[t];
    const int r = threadIdx.x % QR_TILE, g = threadIdx.x / QR_TILE;
    const int j0 = g * CPT;
    const int iend = min(k, j0 + CPT);  // Rinv[i][j] = 0 for i > j
    for (long long r0 = (long long)blockIdx.x * QR_TILE; r0 < n; r0 += (long long)gridDim.x * QR_TILE) {
        __syncthreads();
        qr_load_tile(tile, x, ld, n, r0, k);
        __syncthreads();
        if (j0 < k) {
            double acc[CPT];
#pragma unroll
            for (int b = 0; b < CPT; ++b) acc[b] = 0.0;
            for (int i = 0; i < iend; ++i) {
                const double xv = tile[r][i];
                const double *c = rs + i * QR_MAXK + j0;
#pragma unroll
                for (int b = 0; b < CPT; ++b) acc[b] = fma(xv, c[b], acc[b]);
            }
            if (r0 + r < n) {
#pragma unroll
                for (int b = 0; b < CPT; ++b)
                    if (j0 + b < k) x[(long long)(j0 + b) * ld + r0 + r] = acc[b];
            }
        }
    }
}

template <int KB>
static famg_status launch_gram(cudaStream_t st, int grid, const double *x, int64_t ld, int64_t n, int k, double *partials) {
    gram_partial_kernel<KB><<<grid, QR_THREADS, 0, st>>>(x, ld, n, k, partials);
    return FAMG_OK;
}
template <int KB>
static famg_status launch_rinv(cudaStream_t st, int grid, double *x, int64_t ld, int64_t n, int k, const double *d_rinv) {
    constexpr int SMEM_MAX = (int)(sizeof(double) * (QR_TILE * QR_LDT + QR_MAXK * QR_MAXK));
    static std::atomic<uint64_t> configured{0};  // per template instance, one bit per device (the opt-in is a per-device attribute)
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (!(configured.load(std::memory_order_relaxed) >> (dev & 63) & 1ull)) {
        CUDA_TRY(cudaFuncSetAttribute(apply_rinv_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
        configured.fetch_or(1ull << (dev & 63), std::memory_order_relaxed);
    }
    const size_t smem = sizeof(double) * ((size_t)QR_TILE * QR_LDT + (size_t)k * QR_MAXK);
    apply_rinv_kernel<KB><<<grid, QR_THREADS, smem, st>>>(x, ld, n, k, d_rinv);
    return FAMG_OK;
}

// per-column dot products: grid (nblk, k); partials[c * nblk + blk]
__global__ void __launch_bounds__(QR_THREADS) coldot_partial_kernel(const double *__restrict__ x, long long ldx,
                                                                    const double *__restrict__ y, long long ldy, long long n,
                                                                    double *__restrict__ partials) {
    __shared__ double s_red[QR_THREADS / 32];
    const double *xc = x + (long long)blockIdx.y * ldx, *yc = y + (long long)blockIdx.y * ldy;
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * QR_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * QR_THREADS)
        acc = fma(xc[i], yc[i], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < QR_THREADS / 32; ++w) t += s_red[w];
        partials[(long long)blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
}
__global__ void coldot_reduce_kernel(const double *__restrict__ partials, int nblk, int k, double *__restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= k) return;
    double s = 0.0;
    for (int p = 0; p < nblk; ++p) s += partials[(long long)c * nblk + p];
    out[c] = s;
}

static inline int qr_grid(famg_ctx *ctx, int64_t n) {
    return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, QR_TILE), 3 * (int64_t)ctx->num_sms));  // 3 CTAs/SM (80 registers)
}

// one CholeskyQR pass; d_work holds grid*k*k partials, k*k Gram, k*QR_MAXK Rinv
static famg_status cholqr_pass(famg_ctx *ctx, double *x, int64_t ld, int64_t n, int k, double *d_work, int grid) {
    const int kk = k * k;
    double *partials = d_work, *d_g = d_work + (size_t)grid * kk, *d_rinv = d_g + kk;
    const int kb = (k + 15) / 16;
    switch (kb) {
        case 1: FAMG_TRY(launch_gram<1>(ctx->stream, grid, x, ld, n, k, partials)); break;
        case 2: FAMG_TRY(launch_gram<2>(ctx->stream, grid, x, ld, n, k, partials)); break;
        case 3: FAMG_TRY(launch_gram<3>(ctx->stream, grid, x, ld, n, k, partials)); break;
        default: FAMG_TRY(launch_gram<4>(ctx->stream, grid, x, ld, n, k, partials)); break;
    }
    sum_partials_kernel<<<(kk + 255) / 256, 256, 0, ctx->stream>>>(partials, grid, kk, d_g);
    count_launch(ctx, 2);
    KERNEL_CHECK();
    std::vector<double> g((size_t)kk), l((size_t)kk, 0.0), linv((size_t)kk, 0.0), rinv((size_t)k * QR_MAXK, 0.0);
    CUDA_TRY(cudaMemcpyAsync(g.data(), d_g, sizeof(double) * kk, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    // G = L L^T (lower), R = L^T
    for (int j = 0; j < k; ++j) {
        double d = g[(size_t)j + (size_t)j * k];
        for (int p = 0; p < j; ++p) d -= l[(size_t)j + (size_t)p * k] * l[(size_t)j + (size_t)p * k];
        if (!(d > 0.0)) FAMG_FAIL(FAMG_ERR_NUMERIC, "thin_q: column %d is linearly dependent (Gram pivot %g)", j, d);
        d = sqrt(d);
        l[(size_t)j + (size_t)j * k] = d;
        for (int i = j + 1; i < k; ++i) {
            double s = g[(size_t)i + (size_t)j * k];
            for (int p = 0; p < j; ++p) s -= l[(size_t)i + (size_t)p * k] * l[(size_t)j + (size_t)p * k];
            l[(size_t)i + (size_t)j * k] = s / d;
        }
    }
    // Linv (lower) by forward substitution; Rinv = Linv^T (upper), stored row-major: rinv[i * QR_MAXK + j]
    for (int c = 0; c < k; ++c) {
        for (int i = c; i < k; ++i) {
            double s = i == c ? 1.0 : 0.0;
            for (int p = c; p < i; ++p) s -= l[(size_t)i + (size_t)p * k] * linv[(size_t)p + (size_t)c * k];
            linv[(size_t)i + (size_t)c * k] = s / l[(size_t)i + (size_t)i * k];
        }
    }
    for (int i = 0; i < k; ++i)
        for (int j = i; j < k; ++j) rinv[(size_t)i * QR_MAXK + j] = linv[(size_t)j + (size_t)i * k];
    CUDA_TRY(cudaMemcpyAsync(d_rinv, rinv.data(), sizeof(double) * k * QR_MAXK, cudaMemcpyHostToDevice, ctx->stream));
    switch (kb) {
        case 1: FAMG_TRY(launch_rinv<1>(ctx->stream, grid, x, ld, n, k, d_rinv)); break;
        case 2: FAMG_TRY(launch_rinv<2>(ctx->stream, grid, x, ld, n, k, d_rinv)); break;
        case 3: FAMG_TRY(launch_rinv<3>(ctx->stream, grid, x, ld, n, k, d_rinv)); break;
        default: FAMG_TRY(launch_rinv<4>(ctx->stream, grid, x, ld, n, k, d_rinv)); break;
    }
    count_launch(ctx);
    KERNEL_CHECK();
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // rinv is a host temporary
    return FAMG_OK;
}

static size_t qr_work_bytes(int grid, int k) { return sizeof(double) * ((size_t)grid * k * k + (size_t)k * k + (size_t)k * QR_MAXK); }

static famg_status thin_q_dev(famg_ctx *ctx, double *x, int64_t ld, int64_t n, int k, double *work, int grid) {
    FAMG_TRY(cholqr_pass(ctx, x, ld, n, k, work, grid));
    return cholqr_pass(ctx, x, ld, n, k, work, grid);  // CholeskyQR2
}

static famg_status check_block(const famg_vec *x) {
    if (!x) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    if (x->ncols > QR_MAXK) FAMG_FAIL(FAMG_ERR_UNSUPPORTED, "thin_q_dev supports at most %d columns (got %lld)", QR_MAXK, (long long)x->ncols);
    if (x->nrows < x->ncols) FAMG_FAIL(FAMG_ERR_INVALID, "thin_q_dev needs nrows >= ncols");
    return FAMG_OK;
}

// out[c] = x[:,c] . y[:,c], host result
static famg_status coldot(famg_ctx *ctx, const double *x, int64_t ldx, const double *y, int64_t ldy, int64_t n, int k, double *host) {
    const int nblk = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, (int64_t)QR_THREADS * 8), 4 * (int64_t)ctx->num_sms));
    void *work = nullptr;
    const size_t bytes = sizeof(double) * ((size_t)nblk * k + (size_t)k);
    FAMG_TRY(pool_alloc(ctx, bytes, &work));
    double *partials = (double *)work, *d_out = partials + (size_t)nblk * k;
    coldot_partial_kernel<<<dim3(nblk, k), QR_THREADS, 0, ctx->stream>>>(x, ldx, y, ldy, n, partials);
    coldot_reduce_kernel<<<(k + 63) / 64, 64, 0, ctx->stream>>>(partials, nblk, k, d_out);
    count_launch(ctx, 2);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(host, d_out, sizeof(double) * k, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    pool_free(ctx, work, bytes);
    if (e != cudaSuccess) FAMG_FAIL(FAMG_ERR_CUDA, "column dot failed: %s", cudaGetErrorString(e));
    return FAMG_OK;
}

// out = x - M^-1 (A x); tmp (n x k) is used only when the preconditioner is not a Diag smoother
static famg_status eprop(const famg_csr *a, int pc_kind, void *precond, const double *x, int64_t ldx, double *out, int64_t ldo, double *tmp,
                         int64_t ldt, int k) {
    famg_ctx *ctx = a->ctx;
    SpmvArgs g; g.a = a; g.x = x; g.ldx = ldx; g.y = out; g.ldy = ldo; g.k = k;
    if (pc_kind == FAMG_PC_SMOOTHER && ((const famg_smoother *)precond)->kind == SM_DIAG) {
        g.epi = EPI_EPROP; g.d = ((const famg_smoother *)precond)->d;
        return spmv_launch(g);
    }
    g.epi = EPI_SPMV;
    FAMG_TRY(spmv_launch(g));                                        // out = A x
    famg_vec vo, vt;
    vec_wrap(ctx, out, a->nrows, k, ldo, &vo);
    vec_wrap(ctx, tmp, a->nrows, k, ldt, &vt);
    FAMG_TRY(pc_apply(pc_kind, precond, &vt, &vo));                  // tmp = M^-1 out
    FAMG_TRY(vec_copy(ctx, out, ldo, x, ldx, a->nrows, k));          // out = x
    return vec_axpby(ctx, out, ldo, tmp, ldt, a->nrows, k, -1.0, 1.0);  // out -= tmp
}

static bool is_diag(int pc_kind, const void *precond) {
    return pc_kind == FAMG_PC_SMOOTHER && ((const famg_smoother *)precond)->kind == SM_DIAG;
}

}  // namespace famg

using namespace famg;

extern "C" {

famg_status famg_thin_q_dev(famg_vec *x) {
    FAMG_TRY(check_block(x));
    const int64_t n = x->nrows;
    const int k = (int)x->ncols;
    if (k == 0 || n == 0) return FAMG_OK;
    famg_ctx *ctx = x->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int grid = qr_grid(ctx, n);
    void *work = nullptr;
    const size_t bytes = qr_work_bytes(grid, k);
    FAMG_TRY(pool_alloc(ctx, bytes, &work));
    famg_status st = thin_q_dev(ctx, x->p, x->ld, n, k, (double *)work, grid);
    pool_free(ctx, work, bytes);
    return st;
}

famg_status famg_vec_coldot(const famg_vec *x, const famg_vec *y, double *out) {
    if (!x || !y || !out) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    if (x->nrows != y->nrows || x->ncols != y->ncols) FAMG_FAIL(FAMG_ERR_INVALID, "coldot shape mismatch");
    if (x->ncols == 0) return FAMG_OK;
    CUDA_TRY(cudaSetDevice(x->ctx->device));
    if (x->nrows == 0) { for (int64_t c = 0; c < x->ncols; ++c) out[c] = 0.0; return FAMG_OK; }
    return coldot(x->ctx, x->p, x->ld, y->p, y->ld, x->nrows, (int)x->ncols, out);
}

famg_status famg_error_propagator_dev(const famg_csr *a, const famg_smoother *s, famg_vec *out, const famg_vec *x) {
    if (!a || !s || !out || !x) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    if (a->nrows != a->ncols || s->n != a->nrows || x->nrows != a->nrows || out->nrows != a->nrows || out->ncols != x->ncols)
        FAMG_FAIL(FAMG_ERR_INVALID, "error propagator shape mismatch");
    if (out->p == x->p) FAMG_FAIL(FAMG_ERR_INVALID, "out aliases x");
    famg_ctx *ctx = a->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (x->ncols == 0 || a->nrows == 0) return FAMG_OK;
    famg_vec *tmp = nullptr;
    if (s->kind != SM_DIAG) FAMG_TRY(famg_vec_create(ctx, x->nrows, x->ncols, &tmp));
    famg_status st = eprop(a, FAMG_PC_SMOOTHER, (void *)s, x->p, x->ld, out->p, out->ld, tmp ? tmp->p : nullptr, tmp ? tmp->ld : 0,
                           (int)x->ncols);
    famg_vec_destroy(tmp);
    return st;
}

famg_status famg_smooth_vector_dev(const famg_csr *a, const famg_smoother *s, int64_t iterations, famg_vec *x, double *cfs) {
    if (!s) FAMG_FAIL(FAMG_ERR_INVALID, "bad argument");
    if (a && s->n != a->nrows) FAMG_FAIL(FAMG_ERR_INVALID, "smooth_vector shape mismatch");
    return famg_smooth_vector_pc_dev(a, FAMG_PC_SMOOTHER, (void *)s, iterations, x, cfs);
}

famg_status famg_smooth_vector_pc_dev(const famg_csr *a, int pc_kind, void *precond, int64_t iterations, famg_vec *x, double *cfs) {
    if (!a || !precond || !x || iterations < 0) FAMG_FAIL(FAMG_ERR_INVALID, "bad argument");
    if (pc_kind != FAMG_PC_SMOOTHER && pc_kind != FAMG_PC_MG && pc_kind != FAMG_PC_COMPOSITE) FAMG_FAIL(FAMG_ERR_INVALID, "unknown preconditioner kind");
    if (a->nrows != a->ncols || x->nrows != a->nrows) FAMG_FAIL(FAMG_ERR_INVALID, "smooth_vector shape mismatch");
    FAMG_TRY(check_block(x));
    const int64_t n = x->nrows;
    const int k = (int)x->ncols;
    if (k == 0 || n == 0) return FAMG_OK;
    famg_ctx *ctx = a->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int grid = qr_grid(ctx, n);
    void *work = nullptr;
    const size_t bytes = qr_work_bytes(grid, k);
    famg_vec *y = nullptr, *t = nullptr;
    FAMG_TRY(pool_alloc(ctx, bytes, &work));
    famg_status st = famg_vec_create(ctx, n, k, &y);
    if (st == FAMG_OK && (!is_diag(pc_kind, precond) || cfs)) st = famg_vec_create(ctx, n, k, &t);
    double *cur = x->p; int64_t ldc = x->ld;
    double *oth = y ? y->p : nullptr; int64_t ldo = y ? y->ld : 0;
    // adaptivity.rs:331 and :346: the random block is orthonormalised twice before the loop
    if (st == FAMG_OK) st = thin_q_dev(ctx, cur, ldc, n, k, (double *)work, grid);
    if (st == FAMG_OK) st = thin_q_dev(ctx, cur, ldc, n, k, (double *)work, grid);
    for (int64_t it = 0; it < iterations && st == FAMG_OK; ++it) {  // :351-354
        st = eprop(a, pc_kind, precond, cur, ldc, oth, ldo, t ? t->p : nullptr, t ? t->ld : 0, k);
        std::swap(cur, oth); std::swap(ldc, ldo);
        if (st == FAMG_OK) st = thin_q_dev(ctx, cur, ldc, n, k, (double *)work, grid);
    }
    if (st == FAMG_OK && cur != x->p) { st = vec_copy(ctx, x->p, x->ld, cur, ldc, n, k); oth = cur; ldo = ldc; }
    if (st == FAMG_OK && cfs) {
        // :365-384 for all columns at once:  ||w||_A = sqrt(w.Aw),  Ev = w - M^-1 A w,  cf = ||Ev||_A / ||w||_A
        std::vector<double> wa((size_t)k), ea((size_t)k);
        SpmvArgs g; g.a = a; g.epi = EPI_SPMV; g.x = x->p; g.ldx = x->ld; g.y = oth; g.ldy = ldo; g.k = k;
        st = spmv_launch(g);                                                                        // oth = A w
        if (st == FAMG_OK) st = coldot(ctx, x->p, x->ld, oth, ldo, n, k, wa.data());
        if (st == FAMG_OK) {
            famg_vec vo;
            vec_wrap(ctx, oth, n, k, ldo, &vo);
            st = pc_apply(pc_kind, precond, t, &vo);                                                // t = M^-1 A w
        }
        if (st == FAMG_OK) st = vec_axpby(ctx, t->p, t->ld, x->p, x->ld, n, k, 1.0, -1.0);          // t = w - t
        if (st == FAMG_OK) { g.x = t->p; g.ldx = t->ld; st = spmv_launch(g); }                      // oth = A Ev
        if (st == FAMG_OK) st = coldot(ctx, t->p, t->ld, oth, ldo, n, k, ea.data());
        for (int c = 0; c < k && st == FAMG_OK; ++c) cfs[c] = sqrt(ea[(size_t)c]) / sqrt(wa[(size_t)c]);
    }
    famg_vec_destroy(y);
    famg_vec_destroy(t);
    pool_free(ctx, work, bytes);
    return st;
}

}  // extern "C"
