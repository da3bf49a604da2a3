// common.cuh -- internal types of libfamg (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/famg.h"

namespace famg {

// ---------------------------------------------------------------- error plumbing
void set_error(const char *fmt, ...);
const char *get_error();

#define FAMG_FAIL(code, ...)          \
    do {                              \
        ::famg::set_error(__VA_ARGS__); \
        return (code);                \
    } while (0)

#define CUDA_TRY(expr)                                                                     \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            ::famg::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,           \
                              cudaGetErrorString(e__));                                    \
            return FAMG_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)

#define FAMG_TRY(expr)                 \
    do {                               \
        famg_status s__ = (expr);      \
        if (s__ != FAMG_OK) return s__; \
    } while (0)

#define KERNEL_CHECK()  CUDA_TRY(cudaGetLastError())

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace famg

// ---------------------------------------------------------------- handles
struct famg_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t comm_stream = nullptr;  // halo exchange overlaps interior rows on this stream
    cudaEvent_t ev_release = nullptr;    // orders operator frees after work queued on comm_stream
    std::atomic<bool> comm_used{false};  // set once anything has been queued on comm_stream
    std::atomic<int64_t> option_epoch{0};  // bumped by famg_ctx_set_option: captured graphs are keyed on it
    // in-kernel timeline (set_option("trace", 1)): [0] = record count, then (tag, globaltimer ns) pairs written by the
    // traced kernels; trace_desc[id] describes launch `id` (ids are baked into captured graphs)
    unsigned long long *d_trace = nullptr;
    std::vector<std::string> trace_desc;
    std::atomic<int64_t> launches{0};
    // small persistent scratch: scalars for dots / norms, pinned host mirror
    double *d_scalars = nullptr;  // 64 doubles
    double *h_scalars = nullptr;  // pinned, 64 doubles
    double *d_partials = nullptr; // per-block partial sums
    int64_t partials_cap = 0;
    double *pcg_ws = nullptr;     // PCG work vectors (r, p, z, q), kept so the V-cycle graph is reused
    int64_t pcg_ws_cap = 0;
    std::mutex mu;
    // SpMV kernel selection (spmv.cu): 1 = one staged chunk per CTA, 2 = persistent TMA pipeline for
    // operators with at least tma_min_rows rows.  Overridable with FAMG_SPMV_VARIANT / FAMG_TMA_MIN_ROWS
    // for A/B measurements.
    int spmv_variant = 2;
    int tma_min_rows = 1 << 17;
    int spmm_cb = 2;  // right-hand sides per row walk for k > 1 (1 | 2); FAMG_SPMM_CB / set_option("spmm_cb")
    int block_build_host = 0;  // 1: block smoother blocks inverted on the host (set_option("block_build_host"); A/B of the device build)
};

struct famg_csr {
    famg_ctx *ctx = nullptr;
    std::atomic<int> refs{1};
    int64_t nrows = 0, ncols = 0, nnz = 0;
    int *row_ptr = nullptr;  // nrows+1 (+pad)
    int *col = nullptr;      // nnz (+16 pad, zero filled)
    double *val = nullptr;   // nnz (+16 pad)
    // SpMV plan from row-length statistics
    int tpr = 1;             // threads per row (power of two, 1..32)
    int max_row_nnz = 0;
    double avg_row_nnz = 0.0;
};

struct famg_vec {
    famg_ctx *ctx = nullptr;
    int64_t nrows = 0, ncols = 0, ld = 0;
    double *p = nullptr;
    bool owns = true;
    size_t bytes = 0;  // size of the pooled allocation behind p
};

enum SmootherKind { SM_DIAG = 0, SM_DENSE_INV = 1, SM_SPARSE_INV = 2 };

struct famg_smoother {
    famg_ctx *ctx = nullptr;
    std::atomic<int> refs{1};
    int kind = SM_DIAG;
    int64_t n = 0;
    double *d = nullptr;      // SM_DIAG: n
    double *inv = nullptr;    // SM_DENSE_INV: n x n column-major explicit inverse
    famg_csr *minv = nullptr; // SM_SPARSE_INV: block-diagonal M^-1 as CSR
};

struct famg_partition {        // device-resident aggregates (aggregates.cu)
    famg_ctx *ctx = nullptr;
    int64_t n_nodes = 0, n_aggs = 0;
    int *agg_ptr = nullptr;    // n_aggs + 1
    int *agg_nodes = nullptr;  // n_nodes, ascending inside an aggregate
};

struct famg_composite {
    famg_ctx *ctx = nullptr;
    famg_csr *a = nullptr;  // retained
    std::vector<std::pair<int, void *>> components;  // (pc_kind, borrowed handle)
};

namespace famg {

// z = M^-1 r for any preconditioner kind (pcg.cu); n x k blocks
famg_status pc_apply(int pc_kind, void *precond, famg_vec *z, const famg_vec *r);

// ---------------------------------------------------------------- allocation helpers
// Long-lived device buffers (smoother diagonals, level workspaces, halo lists): carved out of the current device's
// caching pool like everything else (allocation ordered on the legacy stream, where nothing is ever queued: available at
// once; released with cudaFree, which hands the block back to the pool).  Plain cudaMalloc / cudaFree cost milliseconds
// each on the measured boxes and dozens of them sat in every distributed hierarchy build.
template <typename T>
famg_status dev_alloc(T **p, int64_t count) {
    *p = nullptr;
    size_t bytes = sizeof(T) * (size_t)(count > 0 ? count : 1);
    cudaError_t e = cudaMallocAsync((void **)p, bytes, (cudaStream_t)0);
    if (e == cudaErrorMemoryAllocation) {  // give cached blocks back to the driver and retry once
        cudaGetLastError();
        int dev = 0;
        cudaMemPool_t mp;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&mp, dev) == cudaSuccess) { cudaDeviceSynchronize(); cudaMemPoolTrimTo(mp, 0); }
        e = cudaMallocAsync((void **)p, bytes, (cudaStream_t)0);
    }
    if (e != cudaSuccess) {
        *p = nullptr;
        set_error("device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? FAMG_ERR_ALLOC : FAMG_ERR_CUDA;
    }
    return FAMG_OK;
}
// memory that is exported with CUDA IPC (the peer-memory arenas) has to come from cudaMalloc
template <typename T>
famg_status dev_alloc_ipc(T **p, int64_t count) {
    *p = nullptr;
    size_t bytes = sizeof(T) * (size_t)(count > 0 ? count : 1);
    cudaError_t e = cudaMalloc((void **)p, bytes);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? FAMG_ERR_ALLOC : FAMG_ERR_CUDA;
    }
    return FAMG_OK;
}

constexpr int CSR_PAD = 16;  // elements of slack after col/val so aligned 128-bit reads may overrun

famg_status csr_alloc(famg_ctx *ctx, int64_t nrows, int64_t ncols, int64_t nnz, famg_csr **out);
famg_status csr_finalize_plan(famg_csr *a);  // row statistics -> threads-per-row
#ifndef FAMG_TPR_FILL
#define FAMG_TPR_FILL 8.9
#endif
// threads per row from the average row length (the rule csr_finalize_plan applies).  The slabs of a
// distributed operator use the GLOBAL average, so that every rank walks its rows with the same lane
// split -- and therefore the same summation order -- as the undistributed operator.
inline int tpr_for_avg(double avg_row_nnz) {
    int tpr = 1;
    while (tpr < 32 && avg_row_nnz > FAMG_TPR_FILL * tpr) tpr <<= 1;
    return tpr;
}
void csr_release(famg_csr *a);
void smoother_release(famg_smoother *s);
famg_status vec_wrap(famg_ctx *ctx, double *p, int64_t nrows, int64_t ncols, int64_t ld, famg_vec *out);

famg_status ensure_partials(famg_ctx *ctx, int64_t count);
famg_status pool_alloc(famg_ctx *ctx, size_t bytes, void **p);
void pool_free(famg_ctx *ctx, void *p, size_t bytes);
void pool_trim(famg_ctx *ctx);

inline void count_launch(famg_ctx *ctx, int n = 1) { ctx->launches.fetch_add(n, std::memory_order_relaxed); }

// ---------------------------------------------------------------- setup phase timer (FAMG_SETUP_TRACE=1 -> stderr)
struct PhaseTimer {
    famg_ctx *ctx; const char *what; bool on; double t0 = 0;
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    static bool enabled() { static const bool e = getenv("FAMG_SETUP_TRACE") != nullptr; return e; }
    PhaseTimer(famg_ctx *c, const char *w) : ctx(c), what(w), on(enabled()) { if (on) { cudaStreamSynchronize(ctx->stream); t0 = now(); } }
    ~PhaseTimer() { if (on) { cudaStreamSynchronize(ctx->stream); fprintf(stderr, "[setup]   %-34s %9.3f ms\n", what, now() - t0); } }
};

// ---------------------------------------------------------------- in-kernel timeline
constexpr unsigned long long TRACE_CAP = 1ull << 18;  // records
enum TraceKind { TR_BEGIN = 1, TR_END = 2, TR_SIG_OK = 3, TR_PACKED = 4, TR_FLAG_OK = 5 };
// new launch id with a description, or -1 when tracing is off
int trace_new_id(famg_ctx *ctx, const char *fmt, ...);
#ifdef __CUDACC__
__device__ __forceinline__ void trace_rec(unsigned long long *tr, unsigned kind, int id) {
    if (!tr) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    const unsigned long long slot = atomicAdd(tr, 1ull);
    if (slot < TRACE_CAP) {
        tr[1 + 2 * slot] = ((unsigned long long)kind << 48) | ((unsigned long long)(unsigned)id << 16) | (unsigned long long)(blockIdx.x & 0xffffu);
        tr[2 + 2 * slot] = t;
    }
}
#endif

// ---------------------------------------------------------------- device primitives (scan.cu)
// exclusive scan of n ints; out[n] = total (out has n+1 entries). in may alias out.
famg_status exclusive_scan_i32(famg_ctx *ctx, const int *in, int *out, int64_t n);

// ---------------------------------------------------------------- SpMV family (spmv.cu)
enum Epi { EPI_SPMV = 0, EPI_RESID = 1, EPI_SMOOTH = 2, EPI_ADD = 3, EPI_SI = 4, EPI_EPROP = 5 };
struct SpmvArgs {
    const famg_csr *a = nullptr;
    int epi = EPI_SPMV;
    const double *x = nullptr; int64_t ldx = 0;  // gather source (ncols rows)
    double *y = nullptr;       int64_t ldy = 0;  // output (nrows rows)
    const double *b = nullptr; int64_t ldb = 0;  // rhs (RESID, SMOOTH)
    const double *d = nullptr;                   // diagonal (SMOOTH, SI, EPROP)
    int k = 1;
    double *dot_partials = nullptr;              // k == 1: per-CTA partial of sum_i x[i]*(A x)[i] (EPI_SPMV) or of
                                                 // sum_i b[i]*x'[i] (EPI_SMOOTH: PCG's r.z from the cycle's last sweep)
    int row_begin = 0, row_end = -1;             // row range (distributed interior/boundary split)
    int row2_begin = 0, row2_end = 0;            // optional second range handled by the same launch
    int reserve_ctas = 0;                        // leave this many CTA slots free (room for the concurrently running exchange / NCCL kernels)
    // producer-side halo exchange: all rows are computed, [sig_hi, nrows) and [0, sig_lo) first; every consumer warp adds
    // 1 to *sig per finished chunk of those rows (8 per chunk of 256 / tpr rows): spmv_signal_target() in total
    unsigned *sig = nullptr; int sig_lo = 0, sig_hi = 0;
    // fused push (with sig): device P2PPlanDev of the plan through which y is consumed and its per-boundary-row push map
    const void *push_plan = nullptr; const int *push_map = nullptr;
    cudaStream_t stream = nullptr;               // default: ctx stream
};
famg_status spmv_launch(const SpmvArgs &args, int *num_ctas = nullptr);
inline unsigned spmv_signal_target(const famg_csr *a, int sig_lo, int sig_hi) {
    const int rows = 256 / a->tpr;
    return 8u * (unsigned)(ceil_div((int64_t)a->nrows - sig_hi, rows) + ceil_div(sig_lo, rows));
}

// ---------------------------------------------------------------- vector ops (vecops.cu)
famg_status vec_scale_rows(famg_ctx *ctx, const double *d, const double *in, int64_t ldi, double *out,
                           int64_t ldo, int64_t n, int k, cudaStream_t st = nullptr);  // out = d .* in
famg_status vec_fill(famg_ctx *ctx, double *p, int64_t ld, int64_t n, int k, double v);
famg_status vec_copy(famg_ctx *ctx, double *dst, int64_t ldd, const double *src, int64_t lds, int64_t n, int k);
// deterministic reductions: result(s) left in ctx->d_scalars[slot]
famg_status vec_dot(famg_ctx *ctx, const double *x, const double *y, int64_t n, int slot, cudaStream_t st = nullptr);
famg_status reduce_partials(famg_ctx *ctx, const double *partials, int64_t count, int slot, cudaStream_t st = nullptr);
// PCG fused updates (k = 1); scalars live on the device
// x += alpha p; r -= alpha q; partial ||r||^2 -> slot_rr   with alpha = s[num]/s[den]
famg_status pcg_update_xr(famg_ctx *ctx, double *x, double *r, const double *p, const double *q, int64_t n,
                          int slot_num, int slot_den, int slot_rr);
// the same, leaving the per-CTA partials of ||r||^2 in ctx->d_partials (*num_partials of them) for a fused reduction
famg_status pcg_update_xr_partials(famg_ctx *ctx, double *x, double *r, const double *p, const double *q, int64_t n, int slot_num,
                                   int slot_den, int *num_partials);
// p = z + beta p with beta = s[num]/s[den]
famg_status pcg_update_p(famg_ctx *ctx, double *p, const double *z, int64_t n, int slot_num, int slot_den);
famg_status vec_add_inplace(famg_ctx *ctx, double *x, const double *y, int64_t n);  // x += y
famg_status read_scalars(famg_ctx *ctx, int first, int count, double *host);  // sync D2H
famg_status vec_axpby(famg_ctx *ctx, double *y, int64_t ldy, const double *x, int64_t ldx, int64_t n, int k, double alpha, double beta);

// ---------------------------------------------------------------- dense coarse solve (coarse.cu)
famg_status dense_inverse_from_csr(const famg_csr *a, double **d_inv);  // host Cholesky + inverse, uploaded
famg_status dense_gemv(famg_ctx *ctx, const double *inv, int64_t n, const double *x, int64_t ldx, double *y,
                       int64_t ldy, int k, cudaStream_t st = nullptr);

// ---------------------------------------------------------------- smoother apply (smoothers.cu)
famg_status smoother_apply_dev(const famg_smoother *s, const double *in, int64_t ldi, double *out, int64_t ldo,
                               int k, cudaStream_t st = nullptr);

// Diag smoother entries of the rows of `a` (rectangular row slabs allowed: the diagonal of local row i is
// the entry with column id i; rows need not be sorted).  kind: FAMG_DIAG_L1 | FAMG_DIAG_JACOBI.
famg_status diag_from_rows(const famg_csr *a, int kind, double omega, double *d_out);

// tentative prolongator of a scalar problem with one near-null vector from device aggregates (aggregates.cu);
// nn: n_nodes doubles, coarse_nn: n_aggs doubles, both device
famg_status tentative_p1_dev(const famg_partition *part, const double *nn, famg_csr **p_out, double *coarse_nn);

// sparse products (spgemm.cu)
famg_status spgemm_impl(const famg_csr *a, const famg_csr *b, const famg_csr *p_for_smoothing, double omega, famg_csr **out,
                        int epi_kind = 1);
famg_status transpose_impl(const famg_csr *a, famg_csr **out);

// one-sided Jacobi thin SVD of an m x k column-major block held in u (in: A, out: U); gallery.cu
void host_thin_svd(int64_t m, int64_t k, std::vector<double> &u, std::vector<double> &s, std::vector<double> &v);

// host copies of a device CSR (setup-time helpers)
struct HostCsr {
    int64_t nrows = 0, ncols = 0;
    std::vector<int> row_ptr, col;
    std::vector<double> val;
};
famg_status csr_to_host(const famg_csr *a, HostCsr *h);
famg_status csr_from_host_i32(famg_ctx *ctx, int64_t nrows, int64_t ncols, const int *row_ptr, const int *col,
                              const double *val, famg_csr **out);

}  // namespace famg
