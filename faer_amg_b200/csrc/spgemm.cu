// spgemm.cu -- sparse x sparse products, transpose and the Galerkin triple product.
// Reference: faer's `&SparseRowMat * &SparseRowMat` as used by smoothed_aggregation
// (src/interpolation/mod.rs:828 `&r * &(fine_mat * &p)`, :938 `mat * p`), the transpose
// `p.transpose().to_row_major()` (:824-827) and smooth_interpolation (:927-946).
//
// Contract matched ([faer-recalled]): the output pattern is the *structural* product,
// sorted by column inside each row and never pruned; every output entry is the sum of its
// contributions a_ik * b_kj taken in ascending k, each product and each add rounded separately.
//
// Algorithm: row-wise Gustavson with a hash table per output row.
//   pass 0  ub_i = sum_k nnz(B_k)                 upper bound, bins rows into 4 size classes
//   pass 1  insert the candidate columns into the row's table -> exact nnz(C_i); scan -> row_ptr
//   pass 2  re-insert, compact, rank-sort the distinct columns (sorted pattern, bit-exact with a
//           CPU Gustavson), then accumulate: k runs *sequentially* (ascending), the lanes of the
//           row's thread group take the entries of B_k in parallel and locate their slot by binary
//           search.  Within one k every j is unique, so there are no write conflicts and no atomics
//           on values: the result is deterministic and has the reference's summation order.
//   classes warp/128 and warp/1024 keep table + values in shared memory (one warp per row),
//           cta/8192 uses 128 KB of shared memory (one CTA per row), rows above that use
//           per-CTA tables in global memory (persistent CTAs).
// The prolongator-smoothing epilogue  S_i <- -(w/a_ii) S_i + P_i  is fused into pass 2.
#include "common.cuh"

namespace famg {

constexpr int SG_CLASSES = 4;
constexpr int SG_H0 = 128, SG_H1 = 1024, SG_H2 = 8192;

struct SgMat { const int *rp; const int *col; const double *val; };

struct SgEpilogue {
    int enabled;          // smooth_interpolation
    double omega;
    SgMat p;              // P (same rows as C)
    int *error_flag;      // 1: missing / tiny diagonal, 2: pattern(P) not in pattern(A P)
};

__device__ __forceinline__ unsigned sg_hash(int j, int hmask) { return ((unsigned)j * 2654435761u >> 8) & (unsigned)hmask; }

template <int GROUP>
__device__ __forceinline__ void group_sync() {
    if (GROUP == 32) __syncwarp(); else __syncthreads();
}

// Insert every candidate column of row i of A*B into `table` (size H = hmask+1, cleared here).
// Returns the number of distinct columns (same value on every lane of the group).
template <int GROUP>
__device__ int sg_row_insert(int i, int lane, const SgMat &a, const SgMat &b, int *table, int hmask, int *cnt) {
    for (int t = lane; t <= hmask; t += GROUP) table[t] = -1;
    if (lane == 0) *cnt = 0;
    group_sync<GROUP>();
    const int a0 = a.rp[i], a1 = a.rp[i + 1];
    // 8 lanes walk one row of B; GROUP/8 rows of B in flight
    for (int q = a0 + (lane >> 3); q < a1; q += GROUP >> 3) {
        const int k = a.col[q];
        const int b1 = b.rp[k + 1];
        for (int p = b.rp[k] + (lane & 7); p < b1; p += 8) {
            const int j = b.col[p];
            unsigned h = sg_hash(j, hmask);
            while (true) {
                const int old = atomicCAS(&table[h], -1, j);
                if (old == -1) { atomicAdd(cnt, 1); break; }
                if (old == j) break;
                h = (h + 1) & (unsigned)hmask;
            }
        }
    }
    group_sync<GROUP>();
    const int distinct = *cnt;
    group_sync<GROUP>();  // every lane has read the counter before a caller reuses it
    return distinct;
}

template <int GROUP>
__device__ void sg_row_count(int i, int lane, const SgMat &a, const SgMat &b, int *table, int hmask, int *cnt, int *row_nnz) {
    const int c = sg_row_insert<GROUP>(i, lane, a, b, table, hmask, cnt);
    if (lane == 0) row_nnz[i] = c;
    group_sync<GROUP>();
}

// Pass 2 for one row. table/list: H ints each, vals: H doubles.
template <int GROUP>
__device__ void sg_row_fill(int i, int lane, const SgMat &a, const SgMat &b, int *table, int *list, double *vals, int hmask,
                            int *cnt, const int *c_rp, int *c_col, double *c_val, const SgEpilogue &ep) {
    const int n = sg_row_insert<GROUP>(i, lane, a, b, table, hmask, cnt);
    // compact the distinct columns
    if (lane == 0) *cnt = 0;
    group_sync<GROUP>();
    for (int t = lane; t <= hmask; t += GROUP) {
        const int key = table[t];
        if (key != -1) list[atomicAdd(cnt, 1)] = key;
    }
    group_sync<GROUP>();
    // rank sort (keys distinct) into table[0..n)
    for (int e = lane; e < n; e += GROUP) {
        const int key = list[e];
        int rank = 0;
        for (int f = 0; f < n; ++f) rank += list[f] < key;
        table[rank] = key;
    }
    for (int t = lane; t < n; t += GROUP) vals[t] = 0.0;
    group_sync<GROUP>();
    // numeric: ascending k, lanes over the entries of B_k
    const int a0 = a.rp[i], a1 = a.rp[i + 1];
    for (int q = a0; q < a1; ++q) {
        const int k = a.col[q];
        const double av = a.val[q];
        const int b1 = b.rp[k + 1];
        for (int p = b.rp[k] + lane; p < b1; p += GROUP) {
            const int j = b.col[p];
            int lo = 0, hi = n;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (table[mid] < j) lo = mid + 1; else hi = mid;
            }
            vals[lo] = vals[lo] + av * b.val[p];
        }
        group_sync<GROUP>();
    }
    const int base = c_rp[i];
    if (!ep.enabled) {
        for (int t = lane; t < n; t += GROUP) { c_col[base + t] = table[t]; c_val[base + t] = vals[t]; }
    } else {
        // smooth_interpolation: scalar = w * (1/a_ii); v *= -scalar; then += P_ij where present
        double dv = 0.0; bool found = false;
        {
            int lo = a0, hi = a1;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (a.col[mid] < i) lo = mid + 1; else hi = mid; }
            if (lo < a1 && a.col[lo] == i) { dv = a.val[lo]; found = true; }
        }
        if (lane == 0 && (!found || !(dv > 1e-6))) atomicMax(ep.error_flag, 1);
        const double scalar = ep.omega * (1.0 / dv);
        const int p0 = ep.p.rp[i], p1 = ep.p.rp[i + 1];
        int matched = 0;
        for (int t = lane; t < n; t += GROUP) {
            const int j = table[t];
            double v = vals[t] * -scalar;
            int lo = p0, hi = p1;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (ep.p.col[mid] < j) lo = mid + 1; else hi = mid; }
            if (lo < p1 && ep.p.col[lo] == j) { v = v + ep.p.val[lo]; ++matched; }
            c_col[base + t] = j; c_val[base + t] = v;
        }
        // every entry of P_i must have found a slot (add_assign panics otherwise)
        if (GROUP == 32) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) matched += __shfl_xor_sync(0xffffffffu, matched, o);
            if (lane == 0 && matched != p1 - p0) atomicMax(ep.error_flag, 2);
        } else {
            if (lane == 0) *cnt = 0;
            __syncthreads();
            if (matched) atomicAdd(cnt, matched);
            __syncthreads();
            if (lane == 0 && *cnt != p1 - p0) atomicMax(ep.error_flag, 2);
        }
    }
    group_sync<GROUP>();
}

// ---- kernels ------------------------------------------------------------------------------
__global__ void sg_ub_kernel(SgMat a, SgMat b, int m, int *__restrict__ ub, int *__restrict__ cls, int *__restrict__ class_count,
                             int *__restrict__ max_ub) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    long long s = 0;
    for (int q = a.rp[i]; q < a.rp[i + 1]; ++q) { const int k = a.col[q]; s += b.rp[k + 1] - b.rp[k]; }
    const int u = s > 0x3fffffff ? 0x3fffffff : (int)s;
    ub[i] = u;
    const int c = u <= SG_H0 ? 0 : u <= SG_H1 ? 1 : u <= SG_H2 ? 2 : 3;
    cls[i] = c;
    atomicAdd(&class_count[c], 1);
    if (c == 3) atomicMax(max_ub, u);
}

__global__ void sg_bin_kernel(const int *__restrict__ cls, int m, int *__restrict__ cursor, int *__restrict__ perm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    perm[atomicAdd(&cursor[cls[i]], 1)] = i;
}

struct SgArgs {
    SgMat a, b;
    const int *perm; int count;   // rows of this class
    int *row_nnz;                 // pass 1 output
    const int *c_rp; int *c_col; double *c_val;  // pass 2 output
    SgEpilogue ep;
    int *g_table; int *g_list; double *g_vals; int g_hmask;  // class 3 scratch (per CTA slices)
};

template <int H, int WARPS, bool FILL>
__global__ void __launch_bounds__(WARPS * 32) sg_warp_kernel(SgArgs s) {
    __shared__ int s_table[WARPS][H];
    __shared__ int s_list[FILL ? WARPS : 1][FILL ? H : 1];
    __shared__ double s_vals[FILL ? WARPS : 1][FILL ? H : 1];
    __shared__ int s_cnt[WARPS];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int idx = blockIdx.x * WARPS + w;
    if (idx >= s.count) return;
    const int i = s.perm[idx];
    if (FILL) sg_row_fill<32>(i, lane, s.a, s.b, s_table[w], s_list[w], s_vals[w], H - 1, &s_cnt[w], s.c_rp, s.c_col, s.c_val, s.ep);
    else sg_row_count<32>(i, lane, s.a, s.b, s_table[w], H - 1, &s_cnt[w], s.row_nnz);
}

template <bool FILL>
__global__ void __launch_bounds__(256) sg_cta_kernel(SgArgs s) {
    extern __shared__ __align__(16) unsigned char smem[];
    double *vals = reinterpret_cast<double *>(smem);                 // H2 doubles (FILL only)
    int *table = reinterpret_cast<int *>(smem + (FILL ? sizeof(double) * SG_H2 : 0));
    int *list = table + SG_H2;                                       // FILL only
    __shared__ int s_cnt;
    const int i = s.perm[blockIdx.x];
    if (FILL) sg_row_fill<256>(i, threadIdx.x, s.a, s.b, table, list, vals, SG_H2 - 1, &s_cnt, s.c_rp, s.c_col, s.c_val, s.ep);
    else sg_row_count<256>(i, threadIdx.x, s.a, s.b, table, SG_H2 - 1, &s_cnt, s.row_nnz);
}

template <bool FILL>
__global__ void __launch_bounds__(256) sg_global_kernel(SgArgs s) {
    __shared__ int s_cnt;
    const size_t h = (size_t)s.g_hmask + 1;
    int *table = s.g_table + (size_t)blockIdx.x * h;
    int *list = s.g_list + (size_t)blockIdx.x * h;
    double *vals = s.g_vals + (size_t)blockIdx.x * h;
    for (int idx = blockIdx.x; idx < s.count; idx += gridDim.x) {
        const int i = s.perm[idx];
        if (FILL) sg_row_fill<256>(i, threadIdx.x, s.a, s.b, table, list, vals, s.g_hmask, &s_cnt, s.c_rp, s.c_col, s.c_val, s.ep);
        else sg_row_count<256>(i, threadIdx.x, s.a, s.b, table, s.g_hmask, &s_cnt, s.row_nnz);
        __syncthreads();
    }
}

template <bool FILL>
static famg_status sg_launch_all(famg_ctx *ctx, SgArgs base, const int *perm, const int *h_count, int h_max_ub, int **scratch_i,
                                 double **scratch_d) {
    int off = 0;
    for (int c = 0; c < SG_CLASSES; ++c) {
        const int cnt = h_count[c];
        SgArgs s = base;
        s.perm = perm + off; s.count = cnt;
        off += cnt;
        if (cnt == 0) continue;
        if (c == 0) {
            sg_warp_kernel<SG_H0, 8, FILL><<<(unsigned)ceil_div(cnt, 8), 256, 0, ctx->stream>>>(s);
        } else if (c == 1) {
            sg_warp_kernel<SG_H1, 2, FILL><<<(unsigned)ceil_div(cnt, 2), 64, 0, ctx->stream>>>(s);
        } else if (c == 2) {
            const size_t smem = FILL ? (sizeof(double) + 2 * sizeof(int)) * SG_H2 : sizeof(int) * SG_H2;
            CUDA_TRY(cudaFuncSetAttribute(sg_cta_kernel<FILL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            sg_cta_kernel<FILL><<<(unsigned)cnt, 256, smem, ctx->stream>>>(s);
        } else {
            int h = 1;
            while (h < 2 * h_max_ub) h <<= 1;
            const int grid = std::min(cnt, 2 * ctx->num_sms);
            if (!*scratch_i) {
                FAMG_TRY(dev_alloc(scratch_i, (int64_t)grid * h * 2));
                FAMG_TRY(dev_alloc(scratch_d, (int64_t)grid * h));
            }
            s.g_table = *scratch_i; s.g_list = *scratch_i + (size_t)grid * h; s.g_vals = *scratch_d; s.g_hmask = h - 1;
            sg_global_kernel<FILL><<<grid, 256, 0, ctx->stream>>>(s);
        }
        count_launch(ctx);
        KERNEL_CHECK();
    }
    return FAMG_OK;
}

famg_status spgemm_impl(const famg_csr *a, const famg_csr *b, const famg_csr *p_for_smoothing, double omega, famg_csr **out) {
    *out = nullptr;
    if (a->ncols != b->nrows) FAMG_FAIL(FAMG_ERR_INVALID, "spgemm: inner dimensions differ (%lld vs %lld)", (long long)a->ncols, (long long)b->nrows);
    famg_ctx *ctx = a->ctx;
    const int m = (int)a->nrows;
    SgMat A{a->row_ptr, a->col, a->val}, B{b->row_ptr, b->col, b->val};
    int *ub = nullptr, *cls = nullptr, *perm = nullptr, *counters = nullptr, *row_nnz = nullptr, *scratch_i = nullptr;
    double *scratch_d = nullptr;
    famg_csr *c = nullptr;
    famg_status st = FAMG_OK;
    auto cleanup = [&]() {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(ub); cudaFree(cls); cudaFree(perm); cudaFree(counters); cudaFree(row_nnz); cudaFree(scratch_i); cudaFree(scratch_d);
    };
#define SG_TRY(expr) do { st = (expr); if (st != FAMG_OK) { cleanup(); if (c) csr_release(c); return st; } } while (0)
    SG_TRY(dev_alloc(&ub, m));
    SG_TRY(dev_alloc(&cls, m));
    SG_TRY(dev_alloc(&perm, m));
    SG_TRY(dev_alloc(&counters, 16));
    SG_TRY(dev_alloc(&row_nnz, m + 1));
    cudaMemsetAsync(counters, 0, sizeof(int) * 16, ctx->stream);
    int h_counters[16] = {0};
    if (m > 0) {
        sg_ub_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, ctx->stream>>>(A, B, m, ub, cls, counters, counters + 8);
        count_launch(ctx);
        cudaError_t e = cudaMemcpyAsync(h_counters, counters, sizeof(int) * 16, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { set_error("spgemm: %s", cudaGetErrorString(e)); SG_TRY(FAMG_ERR_CUDA); }
        // cursors = exclusive offsets of the classes
        int h_cursor[SG_CLASSES], acc = 0;
        for (int k = 0; k < SG_CLASSES; ++k) { h_cursor[k] = acc; acc += h_counters[k]; }
        cudaMemcpyAsync(counters + 4, h_cursor, sizeof(int) * SG_CLASSES, cudaMemcpyHostToDevice, ctx->stream);
        sg_bin_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, ctx->stream>>>(cls, m, counters + 4, perm);
        count_launch(ctx);
        cudaStreamSynchronize(ctx->stream);  // h_cursor is a stack buffer
    }
    SgArgs base{};
    base.a = A; base.b = B; base.row_nnz = row_nnz;
    base.ep.enabled = 0;
    SG_TRY((sg_launch_all<false>(ctx, base, perm, h_counters, h_counters[8], &scratch_i, &scratch_d)));
    // row_ptr = exclusive scan of the exact row counts
    {
        int *rp = nullptr;
        SG_TRY(dev_alloc(&rp, m + 1));
        st = exclusive_scan_i32(ctx, row_nnz, rp, m);
        int total = 0;
        if (st == FAMG_OK) {
            cudaError_t e = cudaMemcpy(&total, rp + m, sizeof(int), cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { set_error("spgemm: %s", cudaGetErrorString(e)); st = FAMG_ERR_CUDA; }
        }
        if (st == FAMG_OK && total < 0) { set_error("spgemm: product has more than 2^31 non-zeros"); st = FAMG_ERR_UNSUPPORTED; }
        if (st == FAMG_OK) st = csr_alloc(ctx, a->nrows, b->ncols, total, &c);
        if (st == FAMG_OK) cudaMemcpyAsync(c->row_ptr, rp, sizeof(int) * (m + 1), cudaMemcpyDeviceToDevice, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        cudaFree(rp);
        SG_TRY(st);
    }
    int *err_flag = counters + 12;
    base.c_rp = c->row_ptr; base.c_col = c->col; base.c_val = c->val;
    if (p_for_smoothing) {
        base.ep.enabled = 1; base.ep.omega = omega;
        base.ep.p = SgMat{p_for_smoothing->row_ptr, p_for_smoothing->col, p_for_smoothing->val};
        base.ep.error_flag = err_flag;
    }
    SG_TRY((sg_launch_all<true>(ctx, base, perm, h_counters, h_counters[8], &scratch_i, &scratch_d)));
    int h_err = 0;
    {
        cudaError_t e = cudaMemcpyAsync(&h_err, err_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { set_error("spgemm: %s", cudaGetErrorString(e)); SG_TRY(FAMG_ERR_CUDA); }
    }
    if (h_err == 1) { set_error("smooth_interpolation: diagonal missing or nearly zero"); SG_TRY(FAMG_ERR_NUMERIC); }   // interpolation/mod.rs:933-934
    if (h_err == 2) { set_error("smooth_interpolation: pattern(P) is not contained in pattern(A*P)"); SG_TRY(FAMG_ERR_INVALID); }
    SG_TRY(csr_finalize_plan(c));
    cleanup();
#undef SG_TRY
    *out = c;
    return FAMG_OK;
}

// ---------------------------------------------------------------- transpose
__global__ void tr_count_kernel(const int *__restrict__ col, int nnz, int *__restrict__ counts) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += gridDim.x * blockDim.x) atomicAdd(&counts[col[q]], 1);
}
// scatter (source row, value) of every entry into its column bucket; order inside a bucket is
// arbitrary here and fixed by the per-row sort below
__global__ void tr_scatter_kernel(const int *__restrict__ rp, const int *__restrict__ col, const double *__restrict__ val, int m,
                                  int *__restrict__ cursor, int *__restrict__ t_col, double *__restrict__ t_val) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    for (int q = rp[i]; q < rp[i + 1]; ++q) {
        const int pos = atomicAdd(&cursor[col[q]], 1);
        t_col[pos] = i; t_val[pos] = val[q];
    }
}
// one warp per output row: rank sort by (unique) source row index
__global__ void __launch_bounds__(256) tr_sort_kernel(const int *__restrict__ rp, int nrows, const int *__restrict__ t_col,
                                                      const double *__restrict__ t_val, int *__restrict__ o_col,
                                                      double *__restrict__ o_val) {
    const int row = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= nrows) return;
    const int b = rp[row], e = rp[row + 1];
    for (int q = b + lane; q < e; q += 32) {
        const int key = t_col[q];
        int rank = 0;
        for (int f = b; f < e; ++f) rank += t_col[f] < key;
        o_col[b + rank] = key; o_val[b + rank] = t_val[q];
    }
}

famg_status transpose_impl(const famg_csr *a, famg_csr **out) {
    *out = nullptr;
    famg_ctx *ctx = a->ctx;
    const int m = (int)a->nrows, n = (int)a->ncols, nnz = (int)a->nnz;
    famg_csr *t = nullptr;
    FAMG_TRY(csr_alloc(ctx, n, m, nnz, &t));
    int *counts = nullptr, *cursor = nullptr, *t_col = nullptr; double *t_val = nullptr;
    famg_status st = dev_alloc(&counts, n + 1);
    if (st == FAMG_OK) st = dev_alloc(&cursor, n + 1);
    if (st == FAMG_OK) st = dev_alloc(&t_col, nnz);
    if (st == FAMG_OK) st = dev_alloc(&t_val, nnz);
    if (st == FAMG_OK) {
        cudaMemsetAsync(counts, 0, sizeof(int) * (n + 1), ctx->stream);
        if (nnz) {
            const int grid = (int)std::min<int64_t>(ceil_div(nnz, 256), 32 * ctx->num_sms);
            tr_count_kernel<<<grid, 256, 0, ctx->stream>>>(a->col, nnz, counts);
            count_launch(ctx);
        }
        st = exclusive_scan_i32(ctx, counts, t->row_ptr, n);
    }
    if (st == FAMG_OK && nnz) {
        cudaMemcpyAsync(cursor, t->row_ptr, sizeof(int) * (n + 1), cudaMemcpyDeviceToDevice, ctx->stream);
        tr_scatter_kernel<<<(unsigned)ceil_div(std::max(m, 1), 256), 256, 0, ctx->stream>>>(a->row_ptr, a->col, a->val, m, cursor, t_col, t_val);
        tr_sort_kernel<<<(unsigned)ceil_div((int64_t)std::max(n, 1) * 32, 256), 256, 0, ctx->stream>>>(t->row_ptr, n, t_col, t_val, t->col, t->val);
        count_launch(ctx, 2);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { set_error("transpose: %s", cudaGetErrorString(e)); st = FAMG_ERR_CUDA; }
    }
    cudaStreamSynchronize(ctx->stream);
    cudaFree(counts); cudaFree(cursor); cudaFree(t_col); cudaFree(t_val);
    if (st == FAMG_OK) st = csr_finalize_plan(t);
    if (st != FAMG_OK) { csr_release(t); return st; }
    *out = t;
    return FAMG_OK;
}

}  // namespace famg

using namespace famg;

extern "C" {

famg_status famg_spgemm(const famg_csr *a, const famg_csr *b, famg_csr **out) {
    if (!a || !b || !out) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(a->ctx->device));
    return spgemm_impl(a, b, nullptr, 0.0, out);
}

famg_status famg_transpose(const famg_csr *a, famg_csr **out) {
    if (!a || !out) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(a->ctx->device));
    return transpose_impl(a, out);
}

famg_status famg_smooth_interpolation(const famg_csr *a, const famg_csr *p, double omega, famg_csr **out) {
    if (!a || !p || !out) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    if (a->nrows != a->ncols || p->nrows != a->nrows) FAMG_FAIL(FAMG_ERR_INVALID, "smooth_interpolation shape mismatch");
    CUDA_TRY(cudaSetDevice(a->ctx->device));
    return spgemm_impl(a, p, p, omega, out);
}

famg_status famg_galerkin(const famg_csr *a, const famg_csr *p0, int smoothing_steps, double omega, famg_csr **p_out,
                          famg_csr **r_out, famg_csr **ac_out) {
    if (!a || !p0 || !p_out || !r_out || !ac_out || smoothing_steps < 0) FAMG_FAIL(FAMG_ERR_INVALID, "bad argument");
    *p_out = *r_out = *ac_out = nullptr;
    if (a->nrows != a->ncols || p0->nrows != a->nrows) FAMG_FAIL(FAMG_ERR_INVALID, "galerkin shape mismatch");
    CUDA_TRY(cudaSetDevice(a->ctx->device));
    famg_csr *p = const_cast<famg_csr *>(p0);
    p->refs.fetch_add(1);
    famg_status st = FAMG_OK;
    for (int s = 0; s < smoothing_steps && st == FAMG_OK; ++s) {   // interpolation/mod.rs:812-818
        famg_csr *next = nullptr;
        st = spgemm_impl(a, p, p, omega, &next);
        if (st == FAMG_OK) { csr_release(p); p = next; }
    }
    famg_csr *r = nullptr, *ap = nullptr, *ac = nullptr;
    if (st == FAMG_OK) st = transpose_impl(p, &r);                 // :824-827
    if (st == FAMG_OK) st = spgemm_impl(a, p, nullptr, 0.0, &ap);  // :828 inner
    if (st == FAMG_OK) st = spgemm_impl(r, ap, nullptr, 0.0, &ac); // :828 outer
    if (ap) csr_release(ap);
    if (st != FAMG_OK) { csr_release(p); if (r) csr_release(r); if (ac) csr_release(ac); return st; }
    *p_out = p; *r_out = r; *ac_out = ac;
    return FAMG_OK;
}

}  // extern "C"
