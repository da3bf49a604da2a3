// core.cu -- context, CSR container, device multivectors, scan primitive.
#include <cstdarg>

#include <omp.h>

#include "common.cuh"

namespace famg {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char *get_error() { return g_err; }

int trace_new_id(famg_ctx *ctx, const char *fmt, ...) {
    if (!ctx->d_trace) return -1;
    char buf[256];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->trace_desc.emplace_back(buf);
    return (int)ctx->trace_desc.size() - 1;
}

// ---------------------------------------------------------------- scan
// Three-phase exclusive scan (block sums -> recursive scan -> add offsets). Deterministic.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_kernel(const int *__restrict__ in, int *__restrict__ out,
                                                                 int *__restrict__ tile_sums, int64_t n) {
    __shared__ int s_warp[SCAN_THREADS / 32];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t idx = base + i;
        v[i] = idx < n ? in[idx] : 0;
        sum += v[i];
    }
    // warp inclusive scan of per-thread sums
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < SCAN_THREADS / 32 ? s_warp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        if (lane < SCAN_THREADS / 32) s_warp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    int excl = inc - sum + (warp > 0 ? s_warp[warp - 1] : 0);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t idx = base + i;
        if (idx < n) out[idx] = excl;
        excl += v[i];
    }
    if (threadIdx.x == SCAN_THREADS - 1) tile_sums[blockIdx.x] = excl;  // tile total
}

__global__ void scan_add_kernel(int *__restrict__ out, const int *__restrict__ tile_offs, int64_t n) {
    int64_t idx = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x;
    int off = tile_offs[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        int64_t j = idx + (int64_t)i * SCAN_THREADS;
        if (j < n) out[j] += off;
    }
}

__global__ void scan_set_total_kernel(int *out, const int *tile_offs, int64_t n, int64_t ntiles) {
    // out[n] = total = tile_offs[ntiles] (exclusive scan of tile sums has ntiles+1 entries)
    out[n] = tile_offs[ntiles];
}

famg_status exclusive_scan_i32(famg_ctx *ctx, const int *in, int *out, int64_t n) {
    if (n == 0) {
        CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(int), ctx->stream));
        return FAMG_OK;
    }
    int64_t ntiles = ceil_div(n, SCAN_TILE);
    int *tile_sums = nullptr;
    FAMG_TRY(pool_alloc(ctx, sizeof(int) * (size_t)(ntiles + 1), (void **)&tile_sums));
    scan_tile_kernel<<<(unsigned)ntiles, SCAN_THREADS, 0, ctx->stream>>>(in, out, tile_sums, n);
    count_launch(ctx);
    famg_status st = FAMG_OK;
    if (ntiles == 1) {
        // total is tile_sums[0]; offsets are zero
        cudaMemcpyAsync(out + n, tile_sums, sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream);
    } else {
        st = exclusive_scan_i32(ctx, tile_sums, tile_sums, ntiles);  // in place, writes [ntiles]
        if (st == FAMG_OK) {
            scan_add_kernel<<<(unsigned)ntiles, SCAN_THREADS, 0, ctx->stream>>>(out, tile_sums, n);
            scan_set_total_kernel<<<1, 1, 0, ctx->stream>>>(out, tile_sums, n, ntiles);
            count_launch(ctx, 2);
        }
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);  // callers read the total right away
    pool_free(ctx, tile_sums, 0);
    if (st != FAMG_OK) return st;
    if (e != cudaSuccess) FAMG_FAIL(FAMG_ERR_CUDA, "scan failed: %s", cudaGetErrorString(e));
    KERNEL_CHECK();
    return FAMG_OK;
}

// ---------------------------------------------------------------- CSR helpers
static famg_status stream_alloc(famg_ctx *ctx, void **p, size_t bytes) {
    *p = nullptr;
    cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 1, ctx->stream);
    if (e == cudaErrorMemoryAllocation) {  // give cached blocks back and retry once
        cudaGetLastError();
        pool_trim(ctx);
        e = cudaMallocAsync(p, bytes ? bytes : 1, ctx->stream);
    }
    if (e != cudaSuccess) {
        *p = nullptr;
        set_error("cudaMallocAsync(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? FAMG_ERR_ALLOC : FAMG_ERR_CUDA;
    }
    return FAMG_OK;
}

famg_status csr_alloc(famg_ctx *ctx, int64_t nrows, int64_t ncols, int64_t nnz, famg_csr **out) {
    *out = nullptr;
    if (nrows < 0 || ncols < 0 || nnz < 0) FAMG_FAIL(FAMG_ERR_INVALID, "negative CSR dimension");
    if (nnz >= (int64_t)INT32_MAX - CSR_PAD || ncols >= INT32_MAX || nrows >= INT32_MAX)
        FAMG_FAIL(FAMG_ERR_UNSUPPORTED, "CSR too large for 32-bit device indices (nrows=%lld ncols=%lld nnz=%lld)",
                  (long long)nrows, (long long)ncols, (long long)nnz);
    famg_csr *a = new famg_csr();
    a->ctx = ctx; a->nrows = nrows; a->ncols = ncols; a->nnz = nnz;
    // stream-ordered allocation from the device's caching pool (release threshold raised in
    // famg_ctx_create): a hierarchy build allocates and frees dozens of operators, and plain
    // cudaMalloc / cudaFree both synchronise the device and hand memory back to the driver, which
    // made every other build of the same hierarchy take seconds instead of 0.6 s
    famg_status st = stream_alloc(ctx, (void **)&a->row_ptr, sizeof(int) * (size_t)(nrows + 1 + CSR_PAD));
    if (st == FAMG_OK) st = stream_alloc(ctx, (void **)&a->col, sizeof(int) * (size_t)(nnz + CSR_PAD));
    if (st == FAMG_OK) st = stream_alloc(ctx, (void **)&a->val, sizeof(double) * (size_t)(nnz + CSR_PAD));
    if (st != FAMG_OK) { csr_release(a); return st; }
    // zero the pads so aligned over-reads see valid indices / finite values
    cudaMemsetAsync(a->col + nnz, 0, sizeof(int) * CSR_PAD, ctx->stream);
    cudaMemsetAsync(a->val + nnz, 0, sizeof(double) * CSR_PAD, ctx->stream);
    cudaMemsetAsync(a->row_ptr, 0, sizeof(int) * (nrows + 1 + CSR_PAD), ctx->stream);
    *out = a;
    return FAMG_OK;
}

void csr_release(famg_csr *a) {
    if (!a) return;
    if (a->refs.fetch_sub(1) == 1) {
        // ordered after everything already queued on the context's stream that may read the buffers,
        // and after the communication stream (halo exchange: boundary applies read operators there)
        cudaStream_t st = a->ctx->stream;
        if (a->ctx->ev_release && a->ctx->comm_used.load(std::memory_order_relaxed)) {
            cudaEventRecord(a->ctx->ev_release, a->ctx->comm_stream);
            cudaStreamWaitEvent(st, a->ctx->ev_release, 0);
        }
        if (a->row_ptr) cudaFreeAsync(a->row_ptr, st);
        if (a->col) cudaFreeAsync(a->col, st);
        if (a->val) cudaFreeAsync(a->val, st);
        delete a;
    }
}

__global__ void row_max_kernel(const int *__restrict__ row_ptr, int nrows, int *__restrict__ out_max) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int m = 0;
    for (; i < nrows; i += gridDim.x * blockDim.x) m = max(m, row_ptr[i + 1] - row_ptr[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out_max, m);
}

famg_status csr_finalize_plan(famg_csr *a) {
    famg_ctx *ctx = a->ctx;
    a->avg_row_nnz = a->nrows > 0 ? (double)a->nnz / (double)a->nrows : 0.0;
    int *d_max = nullptr;
    FAMG_TRY(pool_alloc(ctx, 256, (void **)&d_max));  // pooled: a cudaMalloc / cudaFree pair per operator adds up over a build
    cudaMemsetAsync(d_max, 0, sizeof(int), ctx->stream);
    if (a->nrows > 0) {
        int blocks = (int)std::min<int64_t>(ceil_div(a->nrows, 256), 4 * ctx->num_sms);
        row_max_kernel<<<blocks, 256, 0, ctx->stream>>>(a->row_ptr, (int)a->nrows, d_max);
        count_launch(ctx);
    }
    int h_max = 0;
    cudaError_t e = cudaMemcpyAsync(&h_max, d_max, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    pool_free(ctx, d_max, 256);
    if (e != cudaSuccess) FAMG_FAIL(FAMG_ERR_CUDA, "csr plan failed: %s", cudaGetErrorString(e));
    a->max_row_nnz = h_max;
    // threads-per-row: each thread should own <= ~9 staged entries so one chunk of 256/tpr rows
    // fits a pipeline stage of 2304 non-zeros (spmv.cu).  Row-length statistics pick scalar-, sub-warp- or
    // warp-per-row.
    a->tpr = tpr_for_avg(a->avg_row_nnz);
    return FAMG_OK;
}

famg_status csr_to_host(const famg_csr *a, HostCsr *h) {
    h->nrows = a->nrows; h->ncols = a->ncols;
    h->row_ptr.resize(a->nrows + 1); h->col.resize(a->nnz); h->val.resize(a->nnz);
    CUDA_TRY(cudaStreamSynchronize(a->ctx->stream));
    CUDA_TRY(cudaMemcpy(h->row_ptr.data(), a->row_ptr, sizeof(int) * (a->nrows + 1), cudaMemcpyDeviceToHost));
    if (a->nnz) {
        CUDA_TRY(cudaMemcpy(h->col.data(), a->col, sizeof(int) * a->nnz, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(h->val.data(), a->val, sizeof(double) * a->nnz, cudaMemcpyDeviceToHost));
    }
    return FAMG_OK;
}

famg_status csr_from_host_i32(famg_ctx *ctx, int64_t nrows, int64_t ncols, const int *row_ptr, const int *col,
                              const double *val, famg_csr **out) {
    int64_t nnz = row_ptr[nrows];
    famg_csr *a = nullptr;
    FAMG_TRY(csr_alloc(ctx, nrows, ncols, nnz, &a));
    cudaError_t e = cudaMemcpyAsync(a->row_ptr, row_ptr, sizeof(int) * (nrows + 1), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(a->col, col, sizeof(int) * nnz, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(a->val, val, sizeof(double) * nnz, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { csr_release(a); FAMG_FAIL(FAMG_ERR_CUDA, "CSR upload failed: %s", cudaGetErrorString(e)); }
    famg_status st = csr_finalize_plan(a);
    if (st != FAMG_OK) { csr_release(a); return st; }
    *out = a;
    return FAMG_OK;
}

famg_status ensure_partials(famg_ctx *ctx, int64_t count) {
    if (count <= ctx->partials_cap) return FAMG_OK;
    if (ctx->d_partials) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->d_partials); ctx->d_partials = nullptr; }
    int64_t cap = std::max<int64_t>(count, 1 << 16);
    FAMG_TRY(dev_alloc(&ctx->d_partials, cap));
    ctx->partials_cap = cap;
    return FAMG_OK;
}

// Temporaries and multivector storage come from the device's stream-ordered caching pool, like operator storage:
// allocation and release are ordered on the context stream, never synchronise the device and never hand memory back to
// the driver.  (Round 1 kept an exact-size free list in front of cudaMalloc / cudaFree; every new size -- and a
// hierarchy build on row slabs has many -- paid a synchronising driver call, with sporadic stalls of 0.1-1.5 s measured
// inside transposes and scans: profiles/r2_setup_phases.md.)
famg_status pool_alloc(famg_ctx *ctx, size_t bytes, void **p) { return stream_alloc(ctx, p, bytes); }

void pool_free(famg_ctx *ctx, void *p, size_t bytes) {
    (void)bytes;
    if (p) cudaFreeAsync(p, ctx->stream);
}

void pool_trim(famg_ctx *ctx) {
    cudaStreamSynchronize(ctx->stream);
    cudaMemPool_t mp;
    if (cudaDeviceGetDefaultMemPool(&mp, ctx->device) == cudaSuccess) cudaMemPoolTrimTo(mp, 0);
}

famg_status vec_wrap(famg_ctx *ctx, double *p, int64_t nrows, int64_t ncols, int64_t ld, famg_vec *out) {
    out->ctx = ctx; out->p = p; out->nrows = nrows; out->ncols = ncols; out->ld = ld; out->owns = false;
    return FAMG_OK;
}

}  // namespace famg

using namespace famg;

// ==================================================================== C ABI: context / CSR / vec
extern "C" {

const char *famg_last_error(void) { return get_error(); }
const char *famg_version(void) { return "faer-amg-b200 0.1 (sm_100a)"; }

famg_status famg_ctx_create(int device, famg_ctx **out) {
    if (!out) FAMG_FAIL(FAMG_ERR_INVALID, "null out");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        FAMG_FAIL(FAMG_ERR_CUDA, "no CUDA device available (%s); libfamg has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= count) FAMG_FAIL(FAMG_ERR_INVALID, "device %d out of range (count %d)", device, count);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        FAMG_FAIL(FAMG_ERR_CUDA, "device %d is sm_%d%d; libfamg is built for sm_100a only", device, prop.major, prop.minor);
    famg_ctx *ctx = new famg_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    if (const char *v = getenv("FAMG_SPMV_VARIANT")) ctx->spmv_variant = atoi(v) == 1 ? 1 : 2;
    if (const char *v = getenv("FAMG_TMA_MIN_ROWS")) ctx->tma_min_rows = std::max(atoi(v), 1);
    if (const char *v = getenv("FAMG_SPMM_CB")) ctx->spmm_cb = atoi(v) >= 2 ? 2 : 1;
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    {   // keep freed operator storage cached in the device pool instead of returning it to the driver
        cudaMemPool_t mp;
        if (cudaDeviceGetDefaultMemPool(&mp, device) == cudaSuccess) {
            uint64_t threshold = UINT64_MAX;
            cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &threshold);
        }
        cudaGetLastError();
    }
    {   // halo exchanges must not queue behind bulk compute: highest priority for the comm stream
        int lo = 0, hi = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, hi));
    }
    CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_release, cudaEventDisableTiming));
    CUDA_TRY(cudaMalloc((void **)&ctx->d_scalars, 64 * sizeof(double)));
    CUDA_TRY(cudaMemset(ctx->d_scalars, 0, 64 * sizeof(double)));
    CUDA_TRY(cudaMallocHost((void **)&ctx->h_scalars, 64 * sizeof(double)));
    *out = ctx;
    return FAMG_OK;
}

famg_status famg_ctx_destroy(famg_ctx *ctx) {
    if (!ctx) return FAMG_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->comm_stream);
    pool_trim(ctx);
    cudaFree(ctx->d_scalars); cudaFreeHost(ctx->h_scalars); cudaFree(ctx->d_partials); cudaFree(ctx->pcg_ws); cudaFree(ctx->d_trace);
    if (ctx->ev_release) cudaEventDestroy(ctx->ev_release);
    cudaStreamDestroy(ctx->stream); cudaStreamDestroy(ctx->comm_stream);
    delete ctx;
    return FAMG_OK;
}

famg_status famg_ctx_sync(famg_ctx *ctx) {
    if (!ctx) FAMG_FAIL(FAMG_ERR_INVALID, "null ctx");
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return FAMG_OK;
}

famg_status famg_ctx_stream(famg_ctx *ctx, void **stream) {
    if (!ctx || !stream) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    *stream = (void *)ctx->stream;
    return FAMG_OK;
}

famg_status famg_ctx_info(famg_ctx *ctx, int *num_sms, int64_t *mem_free, int64_t *mem_total, char *name, int name_len) {
    if (!ctx) FAMG_FAIL(FAMG_ERR_INVALID, "null ctx");
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (num_sms) *num_sms = ctx->num_sms;
    size_t f = 0, t = 0;
    CUDA_TRY(cudaMemGetInfo(&f, &t));
    if (mem_free) *mem_free = (int64_t)f;
    if (mem_total) *mem_total = (int64_t)t;
    if (name && name_len > 0) {
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, ctx->device));
        snprintf(name, name_len, "%s", prop.name);
    }
    return FAMG_OK;
}

famg_status famg_ctx_set_option(famg_ctx *ctx, const char *key, int64_t value) {
    if (!ctx || !key) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    if (!strcmp(key, "spmv_variant")) {
        if (value != 1 && value != 2) FAMG_FAIL(FAMG_ERR_INVALID, "spmv_variant must be 1 or 2");
        ctx->spmv_variant = (int)value;
    } else if (!strcmp(key, "tma_min_rows")) {
        ctx->tma_min_rows = (int)std::max<int64_t>(value, 1);
    } else if (!strcmp(key, "spmm_cb")) {
        if (value != 1 && value != 2) FAMG_FAIL(FAMG_ERR_INVALID, "spmm_cb must be 1 or 2");
        ctx->spmm_cb = (int)value;
    } else if (!strcmp(key, "block_build_host")) {
        ctx->block_build_host = value != 0;
    } else if (!strcmp(key, "trace")) {
        CUDA_TRY(cudaSetDevice(ctx->device));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (value && !ctx->d_trace) {
            CUDA_TRY(cudaMalloc((void **)&ctx->d_trace, sizeof(unsigned long long) * (1 + 2 * TRACE_CAP)));
        }
        if (ctx->d_trace) CUDA_TRY(cudaMemset(ctx->d_trace, 0, sizeof(unsigned long long)));  // value == 0 / repeated 1: restart the record list
    } else {
        FAMG_FAIL(FAMG_ERR_INVALID, "unknown option '%s'", key);
    }
    ctx->option_epoch.fetch_add(1);  // kernel selection changed: graphs captured before must not be replayed
    return FAMG_OK;
}

// Writes the in-kernel timeline collected since set_option("trace", 1) as text: one line per record,
// "<launch id> <kind> <block> <globaltimer ns> <description of the launch>"; kinds: 1 kernel begin, 2 end, 3 exchange
// kernel saw the producer's boundary rows, 4 its entries are packed into the neighbours, 5 the neighbours' entries arrived.
famg_status famg_ctx_trace_dump(famg_ctx *ctx, const char *path) {
    if (!ctx || !path) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    if (!ctx->d_trace) FAMG_FAIL(FAMG_ERR_INVALID, "tracing is off: famg_ctx_set_option(ctx, \"trace\", 1) first");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaDeviceSynchronize());
    unsigned long long n = 0;
    CUDA_TRY(cudaMemcpy(&n, ctx->d_trace, sizeof(n), cudaMemcpyDeviceToHost));
    n = std::min(n, TRACE_CAP);
    std::vector<unsigned long long> rec(2 * (size_t)n);
    if (n) CUDA_TRY(cudaMemcpy(rec.data(), ctx->d_trace + 1, sizeof(unsigned long long) * 2 * n, cudaMemcpyDeviceToHost));
    FILE *f = fopen(path, "w");
    if (!f) FAMG_FAIL(FAMG_ERR_INVALID, "cannot open %s", path);
    for (unsigned long long i = 0; i < n; ++i) {
        const unsigned long long tag = rec[2 * i];
        const int id = (int)((tag >> 16) & 0xffffffffull);
        const char *d = id >= 0 && id < (int)ctx->trace_desc.size() ? ctx->trace_desc[(size_t)id].c_str() : "?";
        fprintf(f, "%d %d %d %llu %s\n", id, (int)(tag >> 48), (int)(tag & 0xffffull), rec[2 * i + 1], d);
    }
    fclose(f);
    return FAMG_OK;
}

// Pre-sizes the context's stream-ordered pool: one allocation of `bytes`, released straight back to the pool, which keeps
// the physical memory (release threshold = max).  Later operator / temporary allocations are carved out of it instead of
// mapping new physical memory in the middle of a hierarchy build (0.2-0.3 ms per MB on the measured boxes).
famg_status famg_ctx_reserve(famg_ctx *ctx, int64_t bytes) {
    if (!ctx || bytes < 0) FAMG_FAIL(FAMG_ERR_INVALID, "bad argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    void *p = nullptr;
    FAMG_TRY(stream_alloc(ctx, &p, (size_t)bytes));
    CUDA_TRY(cudaFreeAsync(p, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return FAMG_OK;
}

famg_status famg_set_num_threads(int n) {
    if (n < 1) FAMG_FAIL(FAMG_ERR_INVALID, "thread count must be positive");
    omp_set_num_threads(n);
    return FAMG_OK;
}

famg_status famg_ctx_launch_count(const famg_ctx *ctx, int64_t *count) {
    if (!ctx || !count) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    *count = ctx->launches.load();
    return FAMG_OK;
}

// ---- CSR
famg_status famg_csr_create(famg_ctx *ctx, int64_t nrows, int64_t ncols, const uint64_t *row_ptr,
                            const uint64_t *col_idx, const double *val, famg_csr **out) {
    if (!ctx || !out || !row_ptr) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    *out = nullptr;
    if (nrows < 0 || ncols < 0) FAMG_FAIL(FAMG_ERR_INVALID, "negative dimension");
    CUDA_TRY(cudaSetDevice(ctx->device));
    uint64_t nnz = row_ptr[nrows];
    if (row_ptr[0] != 0) FAMG_FAIL(FAMG_ERR_INVALID, "row_ptr[0] != 0 (matrix must be compressed)");
    if (nnz >= (uint64_t)INT32_MAX - CSR_PAD || ncols >= INT32_MAX || nrows >= INT32_MAX)
        FAMG_FAIL(FAMG_ERR_UNSUPPORTED, "matrix too large for 32-bit device indices");
    if (nnz && (!col_idx || !val)) FAMG_FAIL(FAMG_ERR_INVALID, "null col_idx/val");
    std::vector<int> rp((size_t)nrows + 1), ci((size_t)nnz);
    for (int64_t i = 0; i <= nrows; ++i) {
        if (i > 0 && (row_ptr[i] < row_ptr[i - 1] || row_ptr[i] > nnz))
            FAMG_FAIL(FAMG_ERR_INVALID, "row_ptr not monotone at row %lld", (long long)i);
        rp[(size_t)i] = (int)row_ptr[i];
    }
    // usize -> i32 narrowing + validation of the faer CSR invariants, rows in parallel
    int64_t bad_range = -1, bad_order = -1;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nrows; ++i)
        for (uint64_t q = row_ptr[i]; q < row_ptr[i + 1]; ++q) {
            const uint64_t c = col_idx[q];
            if (c >= (uint64_t)ncols) {
#pragma omp atomic write
                bad_range = i;
                continue;
            }
            if (q > row_ptr[i] && col_idx[q - 1] >= c) {
#pragma omp atomic write
                bad_order = i;
            }
            ci[(size_t)q] = (int)c;
        }
    if (bad_range >= 0) FAMG_FAIL(FAMG_ERR_INVALID, "column index out of range in row %lld", (long long)bad_range);
    if (bad_order >= 0) FAMG_FAIL(FAMG_ERR_INVALID, "columns of row %lld are not sorted and unique", (long long)bad_order);
    return csr_from_host_i32(ctx, nrows, ncols, rp.data(), ci.data(), val, out);
}

famg_status famg_csr_create_from_triplets(famg_ctx *ctx, int64_t nrows, int64_t ncols, int64_t nt,
                                          const uint64_t *rows, const uint64_t *cols, const double *vals,
                                          famg_csr **out) {
    if (!ctx || !out) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    *out = nullptr;
    if (nt && (!rows || !cols || !vals)) FAMG_FAIL(FAMG_ERR_INVALID, "null triplet arrays");
    CUDA_TRY(cudaSetDevice(ctx->device));
    // host assembly: stable bucket by row, stable sort by column inside the row, merge duplicates
    std::vector<int64_t> start((size_t)nrows + 1, 0);
    for (int64_t t = 0; t < nt; ++t) {
        if (rows[t] >= (uint64_t)nrows || cols[t] >= (uint64_t)ncols) FAMG_FAIL(FAMG_ERR_INVALID, "triplet %lld out of bounds", (long long)t);
        start[(size_t)rows[t] + 1]++;
    }
    for (int64_t i = 0; i < nrows; ++i) start[(size_t)i + 1] += start[(size_t)i];
    std::vector<int64_t> order((size_t)nt), fill(start.begin(), start.end() - 1);
    for (int64_t t = 0; t < nt; ++t) order[(size_t)fill[(size_t)rows[t]]++] = t;
    std::vector<int> rp((size_t)nrows + 1, 0), ci; std::vector<double> cv;
    ci.reserve((size_t)nt); cv.reserve((size_t)nt);
    std::vector<int64_t> tmp;
    for (int64_t i = 0; i < nrows; ++i) {
        tmp.assign(order.begin() + start[(size_t)i], order.begin() + start[(size_t)i + 1]);
        std::stable_sort(tmp.begin(), tmp.end(), [&](int64_t x, int64_t y) { return cols[x] < cols[y]; });
        size_t row_begin = ci.size();
        for (int64_t t : tmp) {
            if (ci.size() > row_begin && (uint64_t)ci.back() == cols[t]) cv.back() += vals[t];
            else { ci.push_back((int)cols[t]); cv.push_back(vals[t]); }
        }
        if (ci.size() >= (size_t)INT32_MAX - CSR_PAD) FAMG_FAIL(FAMG_ERR_UNSUPPORTED, "too many non-zeros");
        rp[(size_t)i + 1] = (int)ci.size();
    }
    return csr_from_host_i32(ctx, nrows, ncols, rp.data(), ci.data(), cv.data(), out);
}

famg_status famg_csr_retain(famg_csr *a) {
    if (!a) FAMG_FAIL(FAMG_ERR_INVALID, "null csr");
    a->refs.fetch_add(1);
    return FAMG_OK;
}
famg_status famg_csr_destroy(famg_csr *a) {
    if (a) { cudaSetDevice(a->ctx->device); csr_release(a); }
    return FAMG_OK;
}
famg_status famg_csr_dims(const famg_csr *a, int64_t *nrows, int64_t *ncols, int64_t *nnz) {
    if (!a) FAMG_FAIL(FAMG_ERR_INVALID, "null csr");
    if (nrows) *nrows = a->nrows;
    if (ncols) *ncols = a->ncols;
    if (nnz) *nnz = a->nnz;
    return FAMG_OK;
}
famg_status famg_csr_download(const famg_csr *a, uint64_t *row_ptr, uint64_t *col_idx, double *val) {
    if (!a || !row_ptr) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(a->ctx->device));
    HostCsr h;
    FAMG_TRY(csr_to_host(a, &h));
    for (int64_t i = 0; i <= a->nrows; ++i) row_ptr[i] = (uint64_t)h.row_ptr[(size_t)i];
    if (col_idx) for (int64_t q = 0; q < a->nnz; ++q) col_idx[q] = (uint64_t)h.col[(size_t)q];
    if (val && a->nnz) memcpy(val, h.val.data(), sizeof(double) * (size_t)a->nnz);
    return FAMG_OK;
}
famg_status famg_csr_plan(const famg_csr *a, int *threads_per_row, int *rows_per_cta, double *avg_row_nnz, int *max_row_nnz) {
    if (!a) FAMG_FAIL(FAMG_ERR_INVALID, "null csr");
    if (threads_per_row) *threads_per_row = a->tpr;
    if (rows_per_cta) *rows_per_cta = 256 / a->tpr;
    if (avg_row_nnz) *avg_row_nnz = a->avg_row_nnz;
    if (max_row_nnz) *max_row_nnz = a->max_row_nnz;
    return FAMG_OK;
}

__global__ void slab_shift_kernel(const int *__restrict__ src, int *__restrict__ dst, int n, int shift) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i] - shift;
}

famg_status famg_csr_row_slab(const famg_csr *a, int64_t row_begin, int64_t row_end, famg_csr **out) {
    if (!a || !out) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    *out = nullptr;
    if (row_begin < 0 || row_end < row_begin || row_end > a->nrows) FAMG_FAIL(FAMG_ERR_INVALID, "bad row range");
    famg_ctx *ctx = a->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    int ends[2];
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaMemcpy(&ends[0], a->row_ptr + row_begin, sizeof(int), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(&ends[1], a->row_ptr + row_end, sizeof(int), cudaMemcpyDeviceToHost));
    int64_t nnz = ends[1] - ends[0], n = row_end - row_begin;
    famg_csr *s = nullptr;
    FAMG_TRY(csr_alloc(ctx, n, a->ncols, nnz, &s));
    slab_shift_kernel<<<(unsigned)ceil_div(n + 1, 256), 256, 0, ctx->stream>>>(a->row_ptr + row_begin, s->row_ptr, (int)(n + 1), ends[0]);
    count_launch(ctx);
    if (nnz) {
        cudaMemcpyAsync(s->col, a->col + ends[0], sizeof(int) * nnz, cudaMemcpyDeviceToDevice, ctx->stream);
        cudaMemcpyAsync(s->val, a->val + ends[0], sizeof(double) * nnz, cudaMemcpyDeviceToDevice, ctx->stream);
    }
    famg_status st = csr_finalize_plan(s);
    if (st != FAMG_OK) { csr_release(s); return st; }
    *out = s;
    return FAMG_OK;
}

// ---- vec
famg_status famg_vec_create(famg_ctx *ctx, int64_t nrows, int64_t ncols, famg_vec **out) {
    if (!ctx || !out || nrows < 0 || ncols < 0) FAMG_FAIL(FAMG_ERR_INVALID, "bad argument");
    *out = nullptr;
    CUDA_TRY(cudaSetDevice(ctx->device));
    famg_vec *v = new famg_vec();
    v->ctx = ctx; v->nrows = nrows; v->ncols = ncols;
    v->ld = (nrows + 1) & ~(int64_t)1;  // keep columns 16-byte aligned
    v->bytes = sizeof(double) * (size_t)(v->ld * std::max<int64_t>(ncols, 1) + 2);
    famg_status st = pool_alloc(ctx, v->bytes, (void **)&v->p);
    if (st != FAMG_OK) { delete v; return st; }
    cudaMemsetAsync(v->p, 0, v->bytes, ctx->stream);
    *out = v;
    return FAMG_OK;
}
famg_status famg_vec_destroy(famg_vec *v) {
    if (!v) return FAMG_OK;
    cudaSetDevice(v->ctx->device);
    if (v->owns) pool_free(v->ctx, v->p, v->bytes);  // stream-ordered reuse, no synchronisation
    delete v;
    return FAMG_OK;
}
famg_status famg_vec_dims(const famg_vec *v, int64_t *nrows, int64_t *ncols) {
    if (!v) FAMG_FAIL(FAMG_ERR_INVALID, "null vec");
    if (nrows) *nrows = v->nrows;
    if (ncols) *ncols = v->ncols;
    return FAMG_OK;
}
famg_status famg_vec_upload(famg_vec *v, const double *host, int64_t ld) {
    if (!v || !host || ld < v->nrows) FAMG_FAIL(FAMG_ERR_INVALID, "bad argument");
    CUDA_TRY(cudaSetDevice(v->ctx->device));
    if (v->nrows && v->ncols)
        CUDA_TRY(cudaMemcpy2DAsync(v->p, sizeof(double) * v->ld, host, sizeof(double) * ld, sizeof(double) * v->nrows,
                                   (size_t)v->ncols, cudaMemcpyHostToDevice, v->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(v->ctx->stream));
    return FAMG_OK;
}
famg_status famg_vec_download(const famg_vec *v, double *host, int64_t ld) {
    if (!v || !host || ld < v->nrows) FAMG_FAIL(FAMG_ERR_INVALID, "bad argument");
    CUDA_TRY(cudaSetDevice(v->ctx->device));
    if (v->nrows && v->ncols)
        CUDA_TRY(cudaMemcpy2DAsync(host, sizeof(double) * ld, v->p, sizeof(double) * v->ld, sizeof(double) * v->nrows,
                                   (size_t)v->ncols, cudaMemcpyDeviceToHost, v->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(v->ctx->stream));
    return FAMG_OK;
}
famg_status famg_vec_fill(famg_vec *v, double value) {
    if (!v) FAMG_FAIL(FAMG_ERR_INVALID, "null vec");
    CUDA_TRY(cudaSetDevice(v->ctx->device));
    return vec_fill(v->ctx, v->p, v->ld, v->nrows, (int)v->ncols, value);
}
famg_status famg_vec_copy(famg_vec *dst, const famg_vec *src) {
    if (!dst || !src || dst->nrows != src->nrows || dst->ncols != src->ncols) FAMG_FAIL(FAMG_ERR_INVALID, "shape mismatch");
    CUDA_TRY(cudaSetDevice(dst->ctx->device));
    return vec_copy(dst->ctx, dst->p, dst->ld, src->p, src->ld, src->nrows, (int)src->ncols);
}
famg_status famg_vec_axpby(famg_vec *y, double alpha, const famg_vec *x, double beta) {
    if (!y || !x || y->nrows != x->nrows || y->ncols != x->ncols) FAMG_FAIL(FAMG_ERR_INVALID, "shape mismatch");
    CUDA_TRY(cudaSetDevice(y->ctx->device));
    return vec_axpby(y->ctx, y->p, y->ld, x->p, x->ld, y->nrows, (int)y->ncols, alpha, beta);
}
famg_status famg_vec_ptr(const famg_vec *v, void **dev_ptr, int64_t *ld) {
    if (!v) FAMG_FAIL(FAMG_ERR_INVALID, "null vec");
    if (dev_ptr) *dev_ptr = v->p;
    if (ld) *ld = v->ld;
    return FAMG_OK;
}
famg_status famg_vec_norm2(const famg_vec *v, double *out) {
    if (!v || !out) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(v->ctx->device));
    for (int64_t c = 0; c < v->ncols; ++c) {
        FAMG_TRY(vec_dot(v->ctx, v->p + c * v->ld, v->p + c * v->ld, v->nrows, 0));
        double s = 0;
        FAMG_TRY(read_scalars(v->ctx, 0, 1, &s));
        out[c] = sqrt(s);
    }
    return FAMG_OK;
}

}  // extern "C"
