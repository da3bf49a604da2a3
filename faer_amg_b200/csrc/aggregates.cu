// aggregates.cu -- device-resident aggregates and the tentative prolongator built from them.
//
// The north-star keeps aggregation on the host "unless it is shown to be on the critical path".  For the benchmark
// configurations it is: with geometric aggregates the only host work of a level is producing the aggregate lists and
// the per-aggregate normalisation of the near-null vector (interpolation/mod.rs:747-809 for one near-null vector), and
// on one B200 those two host passes were 70-90 ms of a 310 ms hierarchy build at 256^3; in a multi-rank job every rank
// only has cores / ranks threads for them, so the setup time GREW with the number of ranks in the weak-scaling runs
// (profiles/r2_setup_phases.md).  Here both run on the device:
//   * famg_partition: aggregate lists (agg_ptr / agg_nodes, nodes ascending inside an aggregate) in device memory,
//     either uploaded from any host partitioner or generated in place for box aggregates of a lexicographic grid
//     (same aggregates, same order as famg_geometric_partition);
//   * famg_tentative_p_dev: one thread per aggregate walks its nodes in list order -- the order of the host loop, so
//     the sum of squares, its square root and the divisions round identically (products and sums are never fused:
//     -fmad=false) -- and writes P and the coarse near-null entry.  Bit-identical to famg_tentative_p.
#include <cmath>

#include "common.cuh"

namespace famg {

__global__ void geo_count_kernel(int nx, int ny, int nz, int bx, int by, int bz, int cx, int cy, int cz, int *__restrict__ cnt) {
    const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long na = (long long)cx * cy * cz;
    if (a >= na) return;
    const int ax = (int)(a % cx), ay = (int)((a / cx) % cy), az = (int)(a / ((long long)cx * cy));
    const int ex = (ax == cx - 1 ? nx : (ax + 1) * bx) - ax * bx;
    const int ey = (ay == cy - 1 ? ny : (ay + 1) * by) - ay * by;
    const int ez = (az == cz - 1 ? nz : (az + 1) * bz) - az * bz;
    cnt[a] = ex * ey * ez;
}
__global__ void geo_fill_kernel(int nx, int ny, int nz, int bx, int by, int bz, int cx, int cy, int cz, const int *__restrict__ ptr,
                                int *__restrict__ nodes) {
    const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long na = (long long)cx * cy * cz;
    if (a >= na) return;
    const int ax = (int)(a % cx), ay = (int)((a / cx) % cy), az = (int)(a / ((long long)cx * cy));
    const int x1 = ax == cx - 1 ? nx : (ax + 1) * bx, y1 = ay == cy - 1 ? ny : (ay + 1) * by, z1 = az == cz - 1 ? nz : (az + 1) * bz;
    int t = ptr[a];
    for (int z = az * bz; z < z1; ++z)
        for (int y = ay * by; y < y1; ++y)
            for (int x = ax * bx; x < x1; ++x) nodes[t++] = x + nx * (y + ny * z);  // ascending: BTreeSet order
}

// one thread per aggregate; `seen` counts how often a node is listed; flag: 1 empty aggregate, 2 node out of range
__global__ void tentative_p1_kernel(const int *__restrict__ agg_ptr, const int *__restrict__ agg_nodes, int n_aggs, int n_fine,
                                    const double *__restrict__ nn, int *__restrict__ col, double *__restrict__ val,
                                    double *__restrict__ coarse_nn, int *__restrict__ seen, int *__restrict__ flag) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_aggs) return;
    const int b0 = agg_ptr[g], b1 = agg_ptr[g + 1];
    if (b1 <= b0) { atomicOr(flag, 1); return; }
    double s = 0.0;
    for (int t = b0; t < b1; ++t) {
        const int node = agg_nodes[t];
        if (node < 0 || node >= n_fine) { atomicOr(flag, 2); return; }
        const double x = nn[node];
        s += x * x;
    }
    const double sv = sqrt(s);
    coarse_nn[g] = sv * 1.0;
    for (int t = b0; t < b1; ++t) {
        const int node = agg_nodes[t];
        col[node] = g;
        val[node] = sv > 0 ? nn[node] / sv : 0.0;
        atomicAdd(&seen[node], 1);
    }
}
__global__ void iota_check_kernel(int *__restrict__ rp, const int *__restrict__ seen, int n, int *__restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) rp[i] = i;
    if (i < n && seen[i] != 1) atomicOr(flag, 4);
}
__global__ void narrow_u64_kernel(const unsigned long long *__restrict__ in, int *__restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int)in[i];
}
__global__ void widen_i32_kernel(const int *__restrict__ in, unsigned long long *__restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (unsigned long long)in[i];
}

static void partition_free(famg_partition *p) {
    if (!p) return;
    if (p->agg_ptr) cudaFreeAsync(p->agg_ptr, p->ctx->stream);
    if (p->agg_nodes) cudaFreeAsync(p->agg_nodes, p->ctx->stream);
    delete p;
}

// P (n_fine x n_aggs, one entry per row) and coarse_nn (device, n_aggs) from device aggregates and a device near-null column
famg_status tentative_p1_dev(const famg_partition *part, const double *nn, famg_csr **p_out, double *coarse_nn) {
    famg_ctx *ctx = part->ctx;
    const int n = (int)part->n_nodes, na = (int)part->n_aggs;
    famg_csr *p = nullptr;
    FAMG_TRY(csr_alloc(ctx, n, na, n, &p));
    int *seen = nullptr;
    famg_status st = pool_alloc(ctx, sizeof(int) * (size_t)(n + 2), (void **)&seen);
    if (st != FAMG_OK) { csr_release(p); return st; }
    int *flag = seen + n;
    cudaMemsetAsync(seen, 0, sizeof(int) * (size_t)(n + 2), ctx->stream);
    if (na) {
        tentative_p1_kernel<<<(unsigned)ceil_div(na, 128), 128, 0, ctx->stream>>>(part->agg_ptr, part->agg_nodes, na, n, nn, p->col, p->val,
                                                                                coarse_nn, seen, flag);
        count_launch(ctx);
    }
    iota_check_kernel<<<(unsigned)ceil_div(n + 1, 256), 256, 0, ctx->stream>>>(p->row_ptr, seen, n, flag);
    count_launch(ctx);
    int h_flag = 0;
    cudaError_t e = cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    pool_free(ctx, seen, 0);
    if (e != cudaSuccess) { csr_release(p); FAMG_FAIL(FAMG_ERR_CUDA, "tentative_p: %s", cudaGetErrorString(e)); }
    if (h_flag & 1) { csr_release(p); FAMG_FAIL(FAMG_ERR_INVALID, "Agg size of 0 cannot support near-null dimension of 1"); }  // interpolation/mod.rs:757-762
    if (h_flag) { csr_release(p); FAMG_FAIL(FAMG_ERR_INVALID, "invalid partition"); }                                               // :745
    p->avg_row_nnz = n > 0 ? 1.0 : 0.0; p->max_row_nnz = n > 0 ? 1 : 0; p->tpr = 1;  // one entry per row: no statistics pass needed
    *p_out = p;
    return FAMG_OK;
}

}  // namespace famg

using namespace famg;

extern "C" {

famg_status famg_partition_geometric_dev(famg_ctx *ctx, int64_t nx, int64_t ny, int64_t nz, int64_t bx, int64_t by, int64_t bz,
                                         famg_partition **out, int64_t *coarse_dims) {
    if (!ctx || !out || nx < 1 || ny < 1 || nz < 1 || bx < 1 || by < 1 || bz < 1) FAMG_FAIL(FAMG_ERR_INVALID, "geometric_partition: bad argument");
    *out = nullptr;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int64_t cx = std::max<int64_t>(nx / bx, 1), cy = std::max<int64_t>(ny / by, 1), cz = std::max<int64_t>(nz / bz, 1);
    const int64_t n = nx * ny * nz, na = cx * cy * cz;
    if (n >= INT32_MAX) FAMG_FAIL(FAMG_ERR_UNSUPPORTED, "grid too large for 32-bit indices");
    if (coarse_dims) { coarse_dims[0] = cx; coarse_dims[1] = cy; coarse_dims[2] = cz; }
    famg_partition *p = new famg_partition();
    p->ctx = ctx; p->n_nodes = n; p->n_aggs = na;
    int *cnt = nullptr;
    famg_status st = pool_alloc(ctx, sizeof(int) * (size_t)(na + 1), (void **)&p->agg_ptr);
    if (st == FAMG_OK) st = pool_alloc(ctx, sizeof(int) * (size_t)std::max<int64_t>(n, 1), (void **)&p->agg_nodes);
    if (st == FAMG_OK) st = pool_alloc(ctx, sizeof(int) * (size_t)(na + 1), (void **)&cnt);
    if (st == FAMG_OK) {
        const unsigned grid = (unsigned)ceil_div(na, 128);
        geo_count_kernel<<<grid, 128, 0, ctx->stream>>>((int)nx, (int)ny, (int)nz, (int)bx, (int)by, (int)bz, (int)cx, (int)cy, (int)cz, cnt);
        count_launch(ctx);
        st = exclusive_scan_i32(ctx, cnt, p->agg_ptr, na);
        if (st == FAMG_OK) {
            geo_fill_kernel<<<grid, 128, 0, ctx->stream>>>((int)nx, (int)ny, (int)nz, (int)bx, (int)by, (int)bz, (int)cx, (int)cy, (int)cz,
                                                           p->agg_ptr, p->agg_nodes);
            count_launch(ctx);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) { set_error("geometric_partition: %s", cudaGetErrorString(e)); st = FAMG_ERR_CUDA; }
        }
    }
    pool_free(ctx, cnt, 0);
    if (st != FAMG_OK) { partition_free(p); return st; }
    *out = p;
    return FAMG_OK;
}

famg_status famg_partition_upload(famg_ctx *ctx, int64_t n_nodes, int64_t n_aggs, const uint64_t *agg_ptr, const uint64_t *agg_nodes,
                                  famg_partition **out) {
    if (!ctx || !out || !agg_ptr || (n_nodes && !agg_nodes) || n_nodes < 0 || n_aggs < 0) FAMG_FAIL(FAMG_ERR_INVALID, "bad argument");
    *out = nullptr;
    if (n_nodes >= INT32_MAX) FAMG_FAIL(FAMG_ERR_UNSUPPORTED, "partition too large for 32-bit indices");
    if ((int64_t)agg_ptr[n_aggs] != n_nodes) FAMG_FAIL(FAMG_ERR_INVALID, "partition does not cover the fine grid");
    CUDA_TRY(cudaSetDevice(ctx->device));
    famg_partition *p = new famg_partition();
    p->ctx = ctx; p->n_nodes = n_nodes; p->n_aggs = n_aggs;
    unsigned long long *tmp = nullptr;
    const size_t m = (size_t)std::max<int64_t>(std::max(n_nodes, n_aggs + 1), 1);
    famg_status st = pool_alloc(ctx, sizeof(int) * (size_t)(n_aggs + 1), (void **)&p->agg_ptr);
    if (st == FAMG_OK) st = pool_alloc(ctx, sizeof(int) * (size_t)std::max<int64_t>(n_nodes, 1), (void **)&p->agg_nodes);
    if (st == FAMG_OK) st = pool_alloc(ctx, sizeof(unsigned long long) * m, (void **)&tmp);
    if (st == FAMG_OK) {
        cudaMemcpyAsync(tmp, agg_ptr, sizeof(uint64_t) * (size_t)(n_aggs + 1), cudaMemcpyHostToDevice, ctx->stream);
        narrow_u64_kernel<<<(unsigned)ceil_div(n_aggs + 1, 256), 256, 0, ctx->stream>>>(tmp, p->agg_ptr, n_aggs + 1);
        if (n_nodes) {
            cudaMemcpyAsync(tmp, agg_nodes, sizeof(uint64_t) * (size_t)n_nodes, cudaMemcpyHostToDevice, ctx->stream);
            narrow_u64_kernel<<<(unsigned)ceil_div(n_nodes, 256), 256, 0, ctx->stream>>>(tmp, p->agg_nodes, n_nodes);
        }
        count_launch(ctx, 2);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);  // the host arrays are borrowed for the duration of the call
        if (e != cudaSuccess) { set_error("partition upload: %s", cudaGetErrorString(e)); st = FAMG_ERR_CUDA; }
    }
    pool_free(ctx, tmp, 0);
    if (st != FAMG_OK) { partition_free(p); return st; }
    *out = p;
    return FAMG_OK;
}

famg_status famg_partition_dims(const famg_partition *p, int64_t *n_nodes, int64_t *n_aggs) {
    if (!p) FAMG_FAIL(FAMG_ERR_INVALID, "null partition");
    if (n_nodes) *n_nodes = p->n_nodes;
    if (n_aggs) *n_aggs = p->n_aggs;
    return FAMG_OK;
}

famg_status famg_partition_download(const famg_partition *p, uint64_t *agg_ptr, uint64_t *agg_nodes) {
    if (!p || !agg_ptr || (p->n_nodes && !agg_nodes)) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    famg_ctx *ctx = p->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    unsigned long long *tmp = nullptr;
    const size_t m = (size_t)std::max<int64_t>(std::max(p->n_nodes, p->n_aggs + 1), 1);
    FAMG_TRY(pool_alloc(ctx, sizeof(unsigned long long) * m, (void **)&tmp));
    widen_i32_kernel<<<(unsigned)ceil_div(p->n_aggs + 1, 256), 256, 0, ctx->stream>>>(p->agg_ptr, tmp, p->n_aggs + 1);
    cudaMemcpyAsync(agg_ptr, tmp, sizeof(uint64_t) * (size_t)(p->n_aggs + 1), cudaMemcpyDeviceToHost, ctx->stream);
    if (p->n_nodes) {
        widen_i32_kernel<<<(unsigned)ceil_div(p->n_nodes, 256), 256, 0, ctx->stream>>>(p->agg_nodes, tmp, p->n_nodes);
        cudaMemcpyAsync(agg_nodes, tmp, sizeof(uint64_t) * (size_t)p->n_nodes, cudaMemcpyDeviceToHost, ctx->stream);
    }
    count_launch(ctx, 2);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    pool_free(ctx, tmp, 0);
    if (e != cudaSuccess) FAMG_FAIL(FAMG_ERR_CUDA, "partition download: %s", cudaGetErrorString(e));
    return FAMG_OK;
}

famg_status famg_partition_destroy(famg_partition *p) {
    if (!p) return FAMG_OK;
    cudaSetDevice(p->ctx->device);
    partition_free(p);
    return FAMG_OK;
}

famg_status famg_tentative_p_dev(const famg_partition *part, const famg_vec *near_null, famg_csr **p, famg_vec *coarse_nn) {
    if (!part || !near_null || !p || !coarse_nn) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    *p = nullptr;
    if (near_null->ncols != 1 || coarse_nn->ncols != 1)
        FAMG_FAIL(FAMG_ERR_UNSUPPORTED, "the device tentative prolongator handles scalar problems with one near-null vector; use famg_tentative_p");
    if (near_null->nrows != part->n_nodes || coarse_nn->nrows != part->n_aggs) FAMG_FAIL(FAMG_ERR_INVALID, "tentative_p: bad dimensions");
    if (near_null->ctx != part->ctx || coarse_nn->ctx != part->ctx) FAMG_FAIL(FAMG_ERR_INVALID, "operands live on different contexts");
    CUDA_TRY(cudaSetDevice(part->ctx->device));
    return tentative_p1_dev(part, near_null->p, p, coarse_nn->p);
}

}  // extern "C"
