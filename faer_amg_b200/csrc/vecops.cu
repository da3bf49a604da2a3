// vecops.cu -- streaming vector kernels: row scaling (Diag apply), fill/copy, deterministic dot
// products and the fused PCG updates (faer conjugate_gradient's vector algebra, utils.rs:600-609).
// Reductions are two-stage (per-CTA partials, then one CTA) -- no floating-point atomics, so every
// run produces the same bits.
#include "common.cuh"

namespace famg {

constexpr int VT = 256;

static inline int vec_grid(famg_ctx *ctx, int64_t n, int per_thread = 4) {
    int64_t blocks = ceil_div(n, (int64_t)VT * per_thread);
    int64_t cap = (int64_t)ctx->num_sms * 8;
    return (int)std::max<int64_t>(1, std::min(blocks, cap));
}

__global__ void __launch_bounds__(VT) scale_rows_kernel(const double *__restrict__ d, const double *__restrict__ in, long long ldi,
                                                        double *__restrict__ out, long long ldo, long long n, int k) {
    for (long long i = (long long)blockIdx.x * VT + threadIdx.x; i < n; i += (long long)gridDim.x * VT) {
        const double di = d[i];
        for (int c = 0; c < k; ++c) out[c * ldo + i] = di * in[c * ldi + i];
    }
}
famg_status vec_scale_rows(famg_ctx *ctx, const double *d, const double *in, int64_t ldi, double *out, int64_t ldo, int64_t n,
                           int k, cudaStream_t st) {
    if (n == 0 || k == 0) return FAMG_OK;
    scale_rows_kernel<<<vec_grid(ctx, n), VT, 0, st ? st : ctx->stream>>>(d, in, ldi, out, ldo, n, k);
    count_launch(ctx);
    KERNEL_CHECK();
    return FAMG_OK;
}

__global__ void __launch_bounds__(VT) fill_kernel(double *__restrict__ p, long long ld, long long n, int k, double v) {
    for (long long i = (long long)blockIdx.x * VT + threadIdx.x; i < n; i += (long long)gridDim.x * VT)
        for (int c = 0; c < k; ++c) p[c * ld + i] = v;
}
famg_status vec_fill(famg_ctx *ctx, double *p, int64_t ld, int64_t n, int k, double v) {
    if (n == 0 || k == 0) return FAMG_OK;
    fill_kernel<<<vec_grid(ctx, n), VT, 0, ctx->stream>>>(p, ld, n, k, v);
    count_launch(ctx);
    KERNEL_CHECK();
    return FAMG_OK;
}

famg_status vec_copy(famg_ctx *ctx, double *dst, int64_t ldd, const double *src, int64_t lds, int64_t n, int k) {
    if (n == 0 || k == 0) return FAMG_OK;
    CUDA_TRY(cudaMemcpy2DAsync(dst, sizeof(double) * ldd, src, sizeof(double) * lds, sizeof(double) * n, (size_t)k,
                               cudaMemcpyDeviceToDevice, ctx->stream));
    return FAMG_OK;
}

// ---- block reduction helper (fixed tree => deterministic)
__device__ __forceinline__ double block_sum(double v, double *s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
    }
    return t;  // valid on thread 0
}

__global__ void __launch_bounds__(VT) dot_partial_kernel(const double *__restrict__ x, const double *__restrict__ y, long long n,
                                                         double *__restrict__ partials) {
    __shared__ double s_red[VT / 32];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * VT + threadIdx.x; i < n; i += (long long)gridDim.x * VT) acc += x[i] * y[i];
    double t = block_sum(acc, s_red);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

__global__ void __launch_bounds__(1024) reduce_partials_kernel(const double *__restrict__ partials, long long count,
                                                              double *__restrict__ scalars, int slot) {
    __shared__ double s_red[32];
    double acc = 0.0;
    for (long long i = threadIdx.x; i < count; i += 1024) acc += partials[i];
    double t = block_sum(acc, s_red);
    if (threadIdx.x == 0) scalars[slot] = t;
}

famg_status reduce_partials(famg_ctx *ctx, const double *partials, int64_t count, int slot, cudaStream_t st) {
    reduce_partials_kernel<<<1, 1024, 0, st ? st : ctx->stream>>>(partials, count, ctx->d_scalars, slot);
    count_launch(ctx);
    KERNEL_CHECK();
    return FAMG_OK;
}

famg_status vec_dot(famg_ctx *ctx, const double *x, const double *y, int64_t n, int slot, cudaStream_t st) {
    int grid = vec_grid(ctx, n, 8);
    FAMG_TRY(ensure_partials(ctx, grid));
    dot_partial_kernel<<<grid, VT, 0, st ? st : ctx->stream>>>(x, y, n, ctx->d_partials);
    count_launch(ctx);
    KERNEL_CHECK();
    return reduce_partials(ctx, ctx->d_partials, grid, slot, st);
}

// x += alpha p; r -= alpha q; partial ||r||^2
__global__ void __launch_bounds__(VT) pcg_xr_kernel(double *__restrict__ x, double *__restrict__ r, const double *__restrict__ p,
                                                    const double *__restrict__ q, long long n, const double *__restrict__ scalars,
                                                    int slot_num, int slot_den, double *__restrict__ partials) {
    __shared__ double s_red[VT / 32];
    const double alpha = scalars[slot_num] / scalars[slot_den];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * VT + threadIdx.x; i < n; i += (long long)gridDim.x * VT) {
        x[i] = x[i] + alpha * p[i];
        const double ri = r[i] - alpha * q[i];
        r[i] = ri;
        acc += ri * ri;
    }
    double t = block_sum(acc, s_red);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
}
famg_status pcg_update_xr_partials(famg_ctx *ctx, double *x, double *r, const double *p, const double *q, int64_t n, int slot_num,
                                   int slot_den, int *num_partials) {
    int grid = vec_grid(ctx, n, 4);
    FAMG_TRY(ensure_partials(ctx, grid));
    pcg_xr_kernel<<<grid, VT, 0, ctx->stream>>>(x, r, p, q, n, ctx->d_scalars, slot_num, slot_den, ctx->d_partials);
    count_launch(ctx);
    KERNEL_CHECK();
    *num_partials = grid;
    return FAMG_OK;
}
famg_status pcg_update_xr(famg_ctx *ctx, double *x, double *r, const double *p, const double *q, int64_t n, int slot_num,
                          int slot_den, int slot_rr) {
    int grid = 0;
    FAMG_TRY(pcg_update_xr_partials(ctx, x, r, p, q, n, slot_num, slot_den, &grid));
    return reduce_partials(ctx, ctx->d_partials, grid, slot_rr, nullptr);
}

__global__ void __launch_bounds__(VT) pcg_p_kernel(double *__restrict__ p, const double *__restrict__ z, long long n,
                                                   const double *__restrict__ scalars, int slot_num, int slot_den) {
    const double beta = scalars[slot_num] / scalars[slot_den];
    for (long long i = (long long)blockIdx.x * VT + threadIdx.x; i < n; i += (long long)gridDim.x * VT) p[i] = z[i] + beta * p[i];
}
famg_status pcg_update_p(famg_ctx *ctx, double *p, const double *z, int64_t n, int slot_num, int slot_den) {
    pcg_p_kernel<<<vec_grid(ctx, n), VT, 0, ctx->stream>>>(p, z, n, ctx->d_scalars, slot_num, slot_den);
    count_launch(ctx);
    KERNEL_CHECK();
    return FAMG_OK;
}

__global__ void __launch_bounds__(VT) add_inplace_kernel(double *__restrict__ x, const double *__restrict__ y, long long n) {
    for (long long i = (long long)blockIdx.x * VT + threadIdx.x; i < n; i += (long long)gridDim.x * VT) x[i] = x[i] + y[i];
}
famg_status vec_add_inplace(famg_ctx *ctx, double *x, const double *y, int64_t n) {
    if (n == 0) return FAMG_OK;
    add_inplace_kernel<<<vec_grid(ctx, n), VT, 0, ctx->stream>>>(x, y, n);
    count_launch(ctx);
    KERNEL_CHECK();
    return FAMG_OK;
}

__global__ void __launch_bounds__(VT) axpby_kernel(double *__restrict__ y, long long ldy, const double *__restrict__ x, long long ldx,
                                                   long long n, int k, double alpha, double beta) {
    for (long long i = (long long)blockIdx.x * VT + threadIdx.x; i < n; i += (long long)gridDim.x * VT)
        for (int c = 0; c < k; ++c) y[c * ldy + i] = alpha * x[c * ldx + i] + beta * y[c * ldy + i];
}
famg_status vec_axpby(famg_ctx *ctx, double *y, int64_t ldy, const double *x, int64_t ldx, int64_t n, int k, double alpha, double beta) {
    if (n == 0 || k == 0) return FAMG_OK;
    axpby_kernel<<<vec_grid(ctx, n), VT, 0, ctx->stream>>>(y, ldy, x, ldx, n, k, alpha, beta);
    count_launch(ctx);
    KERNEL_CHECK();
    return FAMG_OK;
}

famg_status read_scalars(famg_ctx *ctx, int first, int count, double *host) {
    CUDA_TRY(cudaMemcpyAsync(ctx->h_scalars + first, ctx->d_scalars + first, sizeof(double) * count, cudaMemcpyDeviceToHost,
                             ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < count; ++i) host[i] = ctx->h_scalars[first + i];
    return FAMG_OK;
}

}  // namespace famg
