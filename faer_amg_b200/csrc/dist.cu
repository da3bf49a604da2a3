// dist.cu -- row-partitioned multigrid + PCG across GPUs (one process per GPU, NCCL over NVLink).
//
// The reference has no distributed path at all; its row-block parallelism
// (par_spmm.rs:100-110: independent 8192-row tasks reading a shared x) is the decomposition kept
// here, with x no longer shared:
//   * every level l < L_rep is split into contiguous row slabs (one per rank); vectors are split
//     conformally.  A slab's column indices are renumbered to [owned | ghost]; each operator
//     (A_l, R_l, P_l) owns a halo plan: per-peer receive ranges in the ghost tail and per-peer
//     pack lists, both derived at setup from the replicated global matrix (no setup collective).
//   * an apply packs the boundary entries, posts grouped ncclSend/ncclRecv on the comm stream,
//     runs the ghost-free interior rows on the compute stream meanwhile, then the boundary rows.
//   * levels with fewer than `replicate_below` rows per rank (and always the coarsest) are run
//     replicated on the global operators after one gather of the restricted residual; every rank
//     then already holds the full coarse correction, so nothing is scattered back.
//   * PCG dot products are local deterministic reductions + one ncclAllReduce of a single double.
// NCCL is resolved with dlopen at first use so that single-GPU users never need it and the copy
// already loaded by the host process (e.g. torch's) is shared.
#include <dlfcn.h>

#include <cmath>

#include "dist_internal.cuh"

namespace famg {

NcclApi g_nccl;
static std::mutex g_nccl_mu;

famg_status nccl_load() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.handle) return FAMG_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) FAMG_FAIL(FAMG_ERR_COMM, "cannot load libnccl.so.2: %s", dlerror());
#define LOAD(name)                                                                   \
    g_nccl.name = (decltype(g_nccl.name))dlsym(h, "nccl" #name);                     \
    if (!g_nccl.name) { dlclose(h); FAMG_FAIL(FAMG_ERR_COMM, "libnccl lacks nccl" #name); }
    LOAD(GetUniqueId) LOAD(CommInitRank) LOAD(CommDestroy) LOAD(GetErrorString) LOAD(GroupStart) LOAD(GroupEnd)
    LOAD(Send) LOAD(Recv) LOAD(AllReduce) LOAD(Broadcast) LOAD(AllGather)
#undef LOAD
    g_nccl.handle = h;
    return FAMG_OK;
}

__global__ void pack_kernel(const double *__restrict__ x, const int *__restrict__ idx, int n, double *__restrict__ buf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) buf[i] = x[idx[i]];
}

// ---------------------------------------------------------------- peer-memory exchange kernels
// Multi-CTA pack: every CTA stores a slice of each neighbour's entries into that neighbour's
// receive buffer; the CTA that finishes last publishes the epoch (flags) -- the fine-level halo is
// one xy-plane (0.5-2 MB), too much for a single SM's store throughput on the critical path.
__global__ void __launch_bounds__(256) p2p_pack_kernel(const P2PPlanDev pl, const double *__restrict__ x,
                                                       const int *__restrict__ send_idx) {
    __shared__ bool s_last;
    const unsigned long long e = *pl.epoch + 1ull;  // only the last CTA advances *epoch, after everyone has read it
    const int par = (int)(e & 1ull);
    const int stride = gridDim.x * blockDim.x;
    for (int nb = 0; nb < pl.nnb; ++nb) {
        double *dst = pl.rdst[par][nb];
        const int *idx = send_idx + pl.soff[nb];
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < pl.scnt[nb]; j += stride) dst[j] = x[idx[j]];
    }
    __threadfence_system();  // this thread's peer stores are visible system-wide ...
    __syncthreads();         // ... for every thread of the CTA, before the CTA is counted as done
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(pl.done, 1u);
        s_last = prev + 1u == gridDim.x;
    }
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if (threadIdx.x < pl.nnb) st_release_sys(pl.rflag[threadIdx.x], e);
        if (threadIdx.x == 0) { *pl.done = 0u; *pl.epoch = e; }
    }
}

__global__ void __launch_bounds__(256) p2p_wait_kernel(const P2PPlanDev pl, double *__restrict__ ghost_tail) {
    const unsigned long long e = *pl.epoch;  // advanced by this rank's pack kernel (stream order)
    if (threadIdx.x < pl.nnb) {
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(pl.lflag[threadIdx.x]) < e) {
            if (global_timer_ns() - t0 > 20000000000ull) { atomicExch(pl.err, 1); break; }  // 20 s: peer died
        }
    }
    __syncthreads();
    const double *src = pl.lrecv[e & 1ull];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < pl.nghost; j += gridDim.x * blockDim.x)
        ghost_tail[j] = __ldcg(src + j);  // bypass L1 (peer-written)
}

// sum-all-reduce of `count` (<= 4) doubles in scalars[first..]: one CTA of 32 threads per rank
__global__ void __launch_bounds__(32) p2p_allreduce_kernel(const P2PCollDev c, double *__restrict__ scalars, int first, int count) {
    const int lane = threadIdx.x;
    const unsigned long long e = *c.ar_epoch + 1ull;
    const int par = (int)(e & 1ull);
    __syncwarp();
    if (lane < c.nranks) {
        double *dst = c.ar_rslots[par][lane] + c.rank * P2P_AR_WIDTH;  // lane == rank: my own arena
        for (int w = 0; w < count; ++w) dst[w] = scalars[first + w];
        __threadfence_system();
        st_release_sys(c.ar_rflag[lane], e);
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(c.ar_lflag + lane) < e) {
            if (global_timer_ns() - t0 > 20000000000ull) { atomicExch(c.err, 1); break; }
        }
    }
    __syncwarp();
    if (lane < count) {
        const double *src = c.ar_lslots[par];
        double sum = 0.0;
        for (int r = 0; r < c.nranks; ++r) sum += __ldcg(src + r * P2P_AR_WIDTH + lane);  // rank order: same bits everywhere
        scalars[first + lane] = sum;
    }
    if (lane == 0) *c.ar_epoch = e;
}

// all-gather: every rank stores its `cnt` entries at offset `off` of every rank's gather buffer
__global__ void __launch_bounds__(256) p2p_allgather_push_kernel(const P2PCollDev c, const double *__restrict__ local, int off, int cnt) {
    __shared__ bool s_last;
    const unsigned long long e = *c.ag_epoch + 1ull;
    const int par = (int)(e & 1ull);
    for (int p = 0; p < c.nranks; ++p) {
        double *dst = c.ag_rbuf[par][p] + off;
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += gridDim.x * blockDim.x) dst[j] = local[j];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(c.ag_done, 1u) + 1u == gridDim.x;
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if (threadIdx.x < c.nranks) st_release_sys(c.ag_rflag[threadIdx.x], e);
        if (threadIdx.x == 0) { *c.ag_done = 0u; *c.ag_epoch = e; }
    }
}
__global__ void __launch_bounds__(256) p2p_allgather_wait_kernel(const P2PCollDev c, double *__restrict__ global, int n) {
    const unsigned long long e = *c.ag_epoch;
    if (threadIdx.x < c.nranks) {
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(c.ag_lflag + threadIdx.x) < e) {
            if (global_timer_ns() - t0 > 20000000000ull) { atomicExch(c.err, 1); break; }
        }
    }
    __syncthreads();
    const double *src = c.ag_lbuf[e & 1ull];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) global[j] = __ldcg(src + j);
}

// Unsplit applies: pack and wait+copy in ONE launch (<= 64 co-resident CTAs).  Every CTA stores
// its slice into the neighbours' buffers; the last one to finish publishes the epoch; then every CTA
// acquire-spins on this rank's own flags and moves its slice of the ghosts into the vector tail.
__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// `sig` / `target` (producer-side exchange): the kernel is launched next to the kernel that is still computing x_ext and
// starts packing once that kernel's consumer warps have counted `target` finished boundary chunks into *sig.
// `sd` / `sf`: the vector being exchanged is sd .* sf (the first sweep from a zero guess, v = d .* f), still being written
// by a kernel running next to this one: the entries to send are formed here from the operands (same product, same bits).
__global__ void __launch_bounds__(256) p2p_exchange_kernel(const P2PPlanDev pl, double *__restrict__ x_ext, int nloc,
                                                           const int *__restrict__ send_idx, unsigned *sig, unsigned target,
                                                           const double *__restrict__ sd, const double *__restrict__ sf,
                                                           unsigned long long *trace, int trace_id) {
    __shared__ bool s_last;
    if (threadIdx.x == 0) trace_rec(trace, TR_BEGIN, trace_id);
    if (sig) {
        if (threadIdx.x == 0) {
            const unsigned long long t0 = global_timer_ns();
            while (ld_acquire_gpu_u32(sig) < target) {
                if (global_timer_ns() - t0 > 20000000000ull) { atomicExch(pl.err, 2); break; }
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) trace_rec(trace, TR_SIG_OK, trace_id);
    const unsigned long long e = *pl.epoch + 1ull;
    const int par = (int)(e & 1ull);
    const int stride = gridDim.x * blockDim.x;
    for (int nb = 0; nb < pl.nnb; ++nb) {
        double *dst = pl.rdst[par][nb];
        const int *idx = send_idx + pl.soff[nb];
        if (sd) { for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < pl.scnt[nb]; j += stride) { const int i = idx[j]; dst[j] = sd[i] * sf[i]; } }
        else { for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < pl.scnt[nb]; j += stride) dst[j] = x_ext[idx[j]]; }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(pl.done, 1u) + 1u == gridDim.x;
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if (threadIdx.x < pl.nnb) st_release_sys(pl.rflag[threadIdx.x], e);
        if (threadIdx.x == 0) { *pl.done = 0u; *pl.epoch = e; if (sig) *sig = 0u; }  // every CTA is past its wait on *sig
    }
    if (threadIdx.x == 0) trace_rec(trace, TR_PACKED, trace_id);
    if (threadIdx.x < pl.nnb) {
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(pl.lflag[threadIdx.x]) < e) {
            if (global_timer_ns() - t0 > 20000000000ull) { atomicExch(pl.err, 1); break; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) trace_rec(trace, TR_FLAG_OK, trace_id);
    const double *src = pl.lrecv[par];
    double *tail = x_ext + nloc;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < pl.nghost; j += stride) tail[j] = __ldcg(src + j);
    __syncthreads();
    if (threadIdx.x == 0) trace_rec(trace, TR_END, trace_id);
}

// Receiving half of a fused exchange: the producer kernel on every rank has pushed its boundary rows into the peers' receive
// buffers and published the epoch; wait for the neighbours' flags, move the ghosts into the vector's tail, advance the epoch.
__global__ void __launch_bounds__(256) p2p_recv_kernel(const P2PPlanDev pl, double *__restrict__ tail, unsigned long long *trace, int trace_id) {
    __shared__ bool s_last;
    if (threadIdx.x == 0) trace_rec(trace, TR_BEGIN, trace_id);
    const unsigned long long e = *pl.epoch + 1ull;
    if (threadIdx.x < pl.nnb) {
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(pl.lflag[threadIdx.x]) < e) {
            if (global_timer_ns() - t0 > 20000000000ull) { atomicExch(pl.err, 1); break; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) trace_rec(trace, TR_FLAG_OK, trace_id);
    const double *src = pl.lrecv[e & 1ull];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < pl.nghost; j += gridDim.x * blockDim.x) tail[j] = __ldcg(src + j);
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(pl.done, 1u) + 1u == gridDim.x;  // every CTA has read the epoch by now
    __syncthreads();
    if (s_last && threadIdx.x == 0) { *pl.done = 0u; *pl.epoch = e; }
    if (threadIdx.x == 0) trace_rec(trace, TR_END, trace_id);
}

// v = d .* f (first sweep from a zero guess) with the fused push: the entries the peers need are formed, stored to the
// neighbours and published first, then the whole vector is written.
__global__ void __launch_bounds__(256) scale_push_kernel(const double *__restrict__ d, const double *__restrict__ f, double *__restrict__ v,
                                                         int n, const P2PPlanDev *__restrict__ pl, const int *__restrict__ push_map, int push_lo,
                                                         int push_hi, unsigned *sig) {
    const int nb_rows = push_lo + (n - push_hi);
    const unsigned long long e = *pl->epoch + 1ull;
    for (int mi = blockIdx.x * blockDim.x + threadIdx.x; mi < nb_rows; mi += gridDim.x * blockDim.x) {
        const int m = push_map[mi];
        if (m < 0) continue;
        const int row = mi < push_lo ? mi : push_hi + (mi - push_lo);
        pl->rdst[e & 1ull][m >> PUSH_SLOT_BITS][m & ((1 << PUSH_SLOT_BITS) - 1)] = d[row] * f[row];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned old = atomicAdd(sig, 1u);
        if (old + 1u == gridDim.x) {
            __threadfence_system();
            for (int nb = 0; nb < pl->nnb; ++nb) st_release_sys(pl->rflag[nb], e);
            *sig = 0u;
        }
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = d[i] * f[i];
}

static famg_status p2p_recv(famg_comm *cm, const HaloPlan &h, double *x_ext) {
    famg_ctx *ctx = cm->ctx;
    const int grid = std::max(1, std::min(32, (h.nghost + 2047) / 2048));
    const int tid = ctx->d_trace ? trace_new_id(ctx, "recv ghost=%d ctas=%d", h.nghost, grid) : -1;
    p2p_recv_kernel<<<grid, 256, 0, ctx->stream>>>(h.dev, x_ext + h.nloc, ctx->d_trace, tid);
    count_launch(ctx);
    KERNEL_CHECK();
    return FAMG_OK;
}

static famg_status p2p_exchange(famg_comm *cm, const HaloPlan &h, double *x_ext, cudaStream_t st = nullptr, unsigned *sig = nullptr,
                                unsigned target = 0, int max_ctas = 64, const double *sd = nullptr, const double *sf = nullptr) {
    const int work = std::max(h.total_send, h.nghost);
    const int grid = std::max(1, std::min(max_ctas, (work + 1023) / 1024));  // co-resident with the producer kernel when signalled
    famg_ctx *ctx = cm->ctx;
    const int tid = ctx->d_trace ? trace_new_id(ctx, "exchange send=%d ghost=%d ctas=%d%s%s", h.total_send, h.nghost, grid, sig ? " signalled" : "",
                                                sd ? " scaled-pack" : "") : -1;
    p2p_exchange_kernel<<<grid, 256, 0, st ? st : cm->ctx->stream>>>(h.dev, x_ext, h.nloc, h.d_send_idx, sig, target, sd, sf, ctx->d_trace, tid);
    count_launch(cm->ctx);
    KERNEL_CHECK();
    return FAMG_OK;
}

// sum of `count` per-CTA partials (fixed order) and its sum over the ranks, in one launch: scalars[slot]
__global__ void __launch_bounds__(1024) p2p_reduce_allreduce_kernel(const P2PCollDev c, const double *__restrict__ partials, int count,
                                                                    double *__restrict__ scalars, int slot) {
    __shared__ double s_red[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < count; i += 1024) acc += partials[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    double mine = 0.0;
    for (int w = 0; w < 32; ++w) mine += s_red[w];  // every lane of warp 0 forms the same local sum
    const unsigned long long e = *c.ar_epoch + 1ull;
    const int par = (int)(e & 1ull);
    __syncwarp();
    if (lane < c.nranks) {
        c.ar_rslots[par][lane][c.rank * P2P_AR_WIDTH] = mine;  // lane == rank: my own arena
        __threadfence_system();
        st_release_sys(c.ar_rflag[lane], e);
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(c.ar_lflag + lane) < e) {
            if (global_timer_ns() - t0 > 20000000000ull) { atomicExch(c.err, 1); break; }
        }
    }
    __syncwarp();
    if (lane == 0) {
        const double *src = c.ar_lslots[par];
        double sum = 0.0;
        for (int r = 0; r < c.nranks; ++r) sum += __ldcg(src + r * P2P_AR_WIDTH);  // rank order: same bits everywhere
        scalars[slot] = sum;
        *c.ar_epoch = e;
    }
}

static famg_status p2p_begin(famg_comm *cm, const HaloPlan &h, const double *x_ext) {
    const int grid = std::max(1, std::min(64, (h.total_send + 1023) / 1024));
    p2p_pack_kernel<<<grid, 256, 0, cm->ctx->stream>>>(h.dev, x_ext, h.d_send_idx);
    count_launch(cm->ctx);
    KERNEL_CHECK();
    return FAMG_OK;
}
static famg_status p2p_end(famg_comm *cm, const HaloPlan &h, double *x_ext) {
    const int grid = std::max(1, std::min(64, (h.nghost + 1023) / 1024));
    p2p_wait_kernel<<<grid, 256, 0, cm->ctx->stream>>>(h.dev, x_ext + h.nloc);
    count_launch(cm->ctx);
    KERNEL_CHECK();
    return FAMG_OK;
}

// start the halo exchange of x_ext ([owned | ghost tail]) for `op`; compute stream keeps going
static famg_status halo_begin(famg_comm *cm, const HaloPlan &h, double *x_ext) {
    if (!h.any || cm->nranks == 1) return FAMG_OK;
    famg_ctx *ctx = cm->ctx;
    if (h.total_send) {
        pack_kernel<<<(unsigned)ceil_div(h.total_send, 256), 256, 0, ctx->stream>>>(x_ext, h.d_send_idx, h.total_send, h.d_sendbuf);
        count_launch(ctx);
    }
    CUDA_TRY(cudaEventRecord(cm->ev_packed, ctx->stream));
    CUDA_TRY(cudaStreamWaitEvent(ctx->comm_stream, cm->ev_packed, 0));
    ctx->comm_used.store(true, std::memory_order_relaxed);
    NCCL_TRY(g_nccl.GroupStart());
    for (int p = 0; p < cm->nranks; ++p) {
        if (h.send_cnt[p]) NCCL_TRY(g_nccl.Send(h.d_sendbuf + h.send_off[p], (size_t)h.send_cnt[p], ncclDouble, p, cm->comm, ctx->comm_stream));
        if (h.recv_cnt[p]) NCCL_TRY(g_nccl.Recv(x_ext + h.nloc + h.recv_off[p], (size_t)h.recv_cnt[p], ncclDouble, p, cm->comm, ctx->comm_stream));
    }
    NCCL_TRY(g_nccl.GroupEnd());
    CUDA_TRY(cudaEventRecord(cm->ev_halo, ctx->comm_stream));
    return FAMG_OK;
}
static famg_status halo_end(famg_comm *cm, const HaloPlan &h) {
    if (!h.any || cm->nranks == 1) return FAMG_OK;
    CUDA_TRY(cudaStreamWaitEvent(cm->ctx->stream, cm->ev_halo, 0));
    return FAMG_OK;
}

// y = epi(A_local x_ext ...) with the exchange overlapped by the interior rows.
static famg_status dist_apply(famg_comm *cm, const DistOp &op, int epi, double *x_ext, double *y, const double *b, const double *d,
                              double *dot_partials, int *num_partials) {
    SpmvArgs g; g.a = op.local; g.epi = epi; g.x = x_ext; g.ldx = 0; g.y = y; g.ldy = 0; g.b = b; g.ldb = 0; g.d = d; g.k = 1;
    const int nrows = (int)op.local->nrows;
    int total = 0, n = 0;
    if (num_partials) *num_partials = 0;
    if (!op.halo.any || cm->nranks == 1) {
        g.dot_partials = dot_partials;
        FAMG_TRY(spmv_launch(g, &n));
        if (num_partials) *num_partials = n;
        return FAMG_OK;
    }
    if (op.halo.p2p) {
        // pack -> interior rows -> wait+copy ghosts -> boundary rows, all on the compute stream: the
        // neighbours' stores land in this rank's arena while the interior kernel runs
        // Measured on 2 and 8 B200 (scripts/dist_bench.py): with peer-memory stores the ghosts arrive
        // within a few microseconds, less than what splitting the apply into interior + boundary
        // launches costs (persistent-kernel ramp/tail twice, one more dependency) -- even for the
        // finest level.  The split is kept for operators above FAMG_SPLIT_MIN_ROWS (default: never).
        static const int split_min_rows = getenv("FAMG_SPLIT_MIN_ROWS") ? atoi(getenv("FAMG_SPLIT_MIN_ROWS")) : 0x7fffffff;
        if (nrows < split_min_rows) {
            FAMG_TRY(p2p_exchange(cm, op.halo, x_ext));
            g.dot_partials = dot_partials;
            FAMG_TRY(spmv_launch(g, &n));
            if (num_partials) *num_partials = n;
            return FAMG_OK;
        }
        FAMG_TRY(p2p_begin(cm, op.halo, x_ext));
        if (op.ie > op.ib) {
            g.row_begin = op.ib; g.row_end = op.ie; g.dot_partials = dot_partials ? dot_partials + total : nullptr;
            FAMG_TRY(spmv_launch(g, &n)); total += n;
        }
        FAMG_TRY(p2p_end(cm, op.halo, x_ext));
        if (op.ib > 0 || op.ie < nrows) {
            g.row_begin = 0; g.row_end = op.ib; g.row2_begin = op.ie; g.row2_end = nrows;
            g.dot_partials = dot_partials ? dot_partials + total : nullptr;
            FAMG_TRY(spmv_launch(g, &n)); total += n;
        }
        if (num_partials) *num_partials = total;
        return FAMG_OK;
    }
    FAMG_TRY(halo_begin(cm, op.halo, x_ext));
    if (op.ie > op.ib) {
        // the persistent kernel would otherwise occupy every CTA slot for its whole duration and
        // the NCCL send/recv kernel could not start until it retires: leave a few slots free
        g.row_begin = op.ib; g.row_end = op.ie; g.dot_partials = dot_partials ? dot_partials + total : nullptr;
        g.reserve_ctas = 4 * 8;
        FAMG_TRY(spmv_launch(g, &n)); total += n;
        g.reserve_ctas = 0;
    }
    FAMG_TRY(halo_end(cm, op.halo));
    if (op.ib > 0 || op.ie < nrows) {  // both boundary slabs in one launch
        g.row_begin = 0; g.row_end = op.ib; g.row2_begin = op.ie; g.row2_end = nrows;
        g.dot_partials = dot_partials ? dot_partials + total : nullptr;
        FAMG_TRY(spmv_launch(g, &n)); total += n;
    }
    if (num_partials) *num_partials = total;
    return FAMG_OK;
}

static famg_status allreduce_slots(famg_comm *cm, int first, int count) {
    if (cm->nranks == 1) return FAMG_OK;
    double *p = cm->ctx->d_scalars + first;
    NCCL_TRY(g_nccl.AllReduce(p, p, (size_t)count, ncclDouble, ncclSum, cm->comm, cm->ctx->stream));
    return FAMG_OK;
}

static famg_status dist_allreduce(famg_dist_mg *dm, int first, int count) {
    famg_comm *cm = dm->comm;
    if (cm->nranks == 1) return FAMG_OK;
    if (dm->p2p_coll && count <= P2P_AR_WIDTH) {
        p2p_allreduce_kernel<<<1, 32, 0, cm->ctx->stream>>>(dm->coll, cm->ctx->d_scalars, first, count);
        count_launch(cm->ctx);
        KERNEL_CHECK();
        return FAMG_OK;
    }
    return allreduce_slots(cm, first, count);
}

// scalars[slot] = sum over ranks of the sum of `count` partials in ctx->d_partials
static famg_status dist_reduce_allreduce(famg_dist_mg *dm, int count, int slot) {
    famg_comm *cm = dm->comm;
    famg_ctx *ctx = cm->ctx;
    if (cm->nranks > 1 && dm->p2p_coll) {
        p2p_reduce_allreduce_kernel<<<1, 1024, 0, ctx->stream>>>(dm->coll, ctx->d_partials, count, ctx->d_scalars, slot);
        count_launch(ctx);
        KERNEL_CHECK();
        return FAMG_OK;
    }
    FAMG_TRY(reduce_partials(ctx, ctx->d_partials, count, slot));
    return dist_allreduce(dm, slot, 1);
}

// ---- producer-side exchange (peer-memory mode): the exchange of a vector starts on the communication stream as soon
// as the entries its consumers' peers need are final, and runs while the rest of the vector is still being computed.
// fork point: what the exchange kernel has to wait for on the compute stream
static famg_status exch_fork(famg_dist_mg *dm, cudaEvent_t *ev_b) {
    famg_comm *cm = dm->comm;
    famg_ctx *ctx = cm->ctx;
    if (dm->pending) FAMG_FAIL(FAMG_ERR_INVALID, "internal: two halo exchanges in flight");
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    CUDA_TRY(cudaStreamIsCapturing(ctx->stream, &cs));
    const bool cap = cs == cudaStreamCaptureStatusActive;
    *ev_b = cap ? cm->ev_packed : cm->ev_packed_e;
    dm->pending_ev = cap ? cm->ev_halo : cm->ev_halo_e;
    CUDA_TRY(cudaEventRecord(*ev_b, ctx->stream));
    return FAMG_OK;
}
static famg_status exch_launch(famg_dist_mg *dm, cudaEvent_t ev_b, const HaloPlan &h, double *x_ext, unsigned *sig, unsigned target,
                               const double *sd = nullptr, const double *sf = nullptr) {
    famg_comm *cm = dm->comm;
    famg_ctx *ctx = cm->ctx;
    CUDA_TRY(cudaStreamWaitEvent(ctx->comm_stream, ev_b, 0));
    ctx->comm_used.store(true, std::memory_order_relaxed);
    // next to a running producer the exchange kernel only gets the CTA slots that kernel left free
    FAMG_TRY(p2p_exchange(cm, h, x_ext, ctx->comm_stream, sig, target, (sig || sd) ? std::max(dm->reserve_ctas, 1) : 64, sd, sf));
    CUDA_TRY(cudaEventRecord(dm->pending_ev, ctx->comm_stream));
    dm->pending = true;
    return FAMG_OK;
}
// exchange of a finished vector on the communication stream
static famg_status exch_async(famg_dist_mg *dm, const HaloPlan &h, double *x_ext) {
    if (!h.any || dm->comm->nranks == 1) return FAMG_OK;
    cudaEvent_t ev_b = nullptr;
    FAMG_TRY(exch_fork(dm, &ev_b));
    return exch_launch(dm, ev_b, h, x_ext, nullptr, 0);
}
static famg_status exch_wait(famg_dist_mg *dm) {
    if (!dm->pending) return FAMG_OK;
    CUDA_TRY(cudaStreamWaitEvent(dm->comm->ctx->stream, dm->pending_ev, 0));
    dm->pending = false;
    return FAMG_OK;
}

// y = epi(A_local x_ext ...) where the ghost tail of x_ext was filled by the exchange started by its producer; `next` is
// the halo plan through which y will be consumed (nullptr: nobody reads its ghosts).  The rows whose results the peers
// need run first, then the exchange of y starts on the communication stream while the remaining rows are computed.
static famg_status dist_apply_push(famg_dist_mg *dm, const DistOp &op, int epi, double *x_ext, double *y, const double *b, const double *d,
                                   const HaloPlan *next, double *dot_partials, int *num_partials) {
    SpmvArgs g; g.a = op.local; g.epi = epi; g.x = x_ext; g.ldx = 0; g.y = y; g.ldy = 0; g.b = b; g.ldb = 0; g.d = d; g.k = 1;
    int n = 0;
    if (num_partials) *num_partials = 0;
    FAMG_TRY(exch_wait(dm));
    const bool push = next && next->any && dm->comm->nranks > 1;
    g.dot_partials = dot_partials;
    if (!push) {
        FAMG_TRY(spmv_launch(g, &n));
        if (num_partials) *num_partials = n;
        return FAMG_OK;
    }
    if (dm->overlap_mode == 3 && next->fusable && spmv_signal_target(op.local, next->push_lo, next->push_hi) > 0) {
        // Fused push: the kernel computes the rows the peers need first and stores them straight into the neighbours' receive
        // buffers; the warp finishing the last of them publishes the epoch.  What is left of the exchange is a small
        // kernel after it that finds the neighbours' flags already set and moves the ghosts into the tail.
        g.sig = dm->d_sig; g.sig_lo = next->push_lo; g.sig_hi = next->push_hi;
        g.push_plan = next->d_dev; g.push_map = next->d_push_map;
        FAMG_TRY(spmv_launch(g, &n));
        if (num_partials) *num_partials = n;
        return p2p_recv(dm->comm, *next, y);
    }
    if (dm->overlap_mode != 2) {  // exchange after the whole apply (no overlap)
        FAMG_TRY(spmv_launch(g, &n));
        if (num_partials) *num_partials = n;
        return exch_async(dm, *next, y);
    }
    // One launch computes the rows the peers need first and counts them into *sig; the exchange kernel is launched
    // BEFORE it on the communication stream, waits for the count, and then packs / pushes / waits for the neighbours
    // while this kernel is still busy with the remaining rows.
    // (The producer is launched first: if a tool serialises kernels in launch order the exchange simply runs after it.)
    const int lo = next->push_lo, hi = next->push_hi;
    cudaEvent_t ev_b = nullptr;
    FAMG_TRY(exch_fork(dm, &ev_b));
    g.sig = dm->d_sig; g.sig_lo = lo; g.sig_hi = hi;
    g.reserve_ctas = dm->reserve_ctas;  // the persistent kernel leaves CTA slots free for the exchange kernel next to it
    FAMG_TRY(spmv_launch(g, &n));
    if (num_partials) *num_partials = n;
    return exch_launch(dm, ev_b, *next, y, dm->d_sig, spmv_signal_target(op.local, lo, hi));
}

// gather the owned pieces of a level vector into a replicated one
static famg_status allgather_rows(famg_dist_mg *dm, int level, const double *local, double *global) {
    famg_comm *cm = dm->comm;
    const auto &sp = dm->splits[level];
    if (dm->p2p_coll && cm->nranks > 1 && level == dm->lrep) {
        const int cnt = (int)(sp[cm->rank + 1] - sp[cm->rank]), n = (int)sp[cm->nranks];
        p2p_allgather_push_kernel<<<std::max(1, std::min(32, (cnt + 1023) / 1024)), 256, 0, cm->ctx->stream>>>(dm->coll, local, (int)sp[cm->rank], cnt);
        p2p_allgather_wait_kernel<<<std::max(1, std::min(32, (n + 1023) / 1024)), 256, 0, cm->ctx->stream>>>(dm->coll, global, n);
        count_launch(cm->ctx, 2);
        KERNEL_CHECK();
        return FAMG_OK;
    }
    if (cm->nranks == 1) {
        CUDA_TRY(cudaMemcpyAsync(global, local, sizeof(double) * (sp[1] - sp[0]), cudaMemcpyDeviceToDevice, cm->ctx->stream));
        return FAMG_OK;
    }
    NCCL_TRY(g_nccl.GroupStart());
    for (int p = 0; p < cm->nranks; ++p) {
        const size_t cnt = (size_t)(sp[p + 1] - sp[p]);
        if (!cnt) continue;
        NCCL_TRY(g_nccl.Broadcast(p == cm->rank ? local : global + sp[p], global + sp[p], cnt, ncclDouble, p, cm->comm, cm->ctx->stream));
    }
    NCCL_TRY(g_nccl.GroupEnd());
    return FAMG_OK;
}

// One visit of a distributed level with producer-side exchanges (peer-memory mode); result in va.  `final_plan`: the
// plan through which the caller consumes va (the parent's prolongator), nullptr on the finest level, where the last
// sweep instead leaves the partial sums of f . va (PCG's r.z) in dot_partials.
static famg_status dist_cycle_push(famg_dist_mg *dm, int level, double *va, const double *f, bool zero_guess, const HaloPlan *final_plan,
                                   double *dot_partials, int *dot_count) {
    famg_comm *cm = dm->comm;
    famg_ctx *ctx = cm->ctx;
    famg_mg *gm = dm->global;
    DistLevel &L = dm->lv[(size_t)level];
    const int64_t nloc = L.r1 - L.r0;
    double *cur = va, *oth = L.t;
    const int nu = gm->nu, mu = gm->mu;
    const HaloPlan *pa = &L.A->halo;
    int pre = nu;
    if (zero_guess) {
        if (((nu - 1) + nu) & 1) std::swap(cur, oth);
        if (dm->overlap_mode == 3 && pa->fusable && pa->any && cm->nranks > 1 && nloc > 0) {
            const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(nloc, 1024), 8 * (int64_t)ctx->num_sms));
            scale_push_kernel<<<grid, 256, 0, ctx->stream>>>(L.d, f, cur, (int)nloc, pa->d_dev, pa->d_push_map, pa->push_lo, pa->push_hi, dm->d_sig);
            count_launch(ctx);
            KERNEL_CHECK();
            FAMG_TRY(p2p_recv(cm, *pa, cur));
        } else if (dm->overlap_mode == 2 && pa->any && cm->nranks > 1) {
            // v = d .* f: the exchange kernel forms the entries the peers need from d and f itself, next to the scaling kernel
            cudaEvent_t ev_b = nullptr;
            FAMG_TRY(exch_fork(dm, &ev_b));
            FAMG_TRY(vec_scale_rows(ctx, L.d, f, 0, cur, 0, nloc, 1));
            FAMG_TRY(exch_launch(dm, ev_b, *pa, cur, nullptr, 0, L.d, f));
        } else {
            FAMG_TRY(vec_scale_rows(ctx, L.d, f, 0, cur, 0, nloc, 1));
            FAMG_TRY(exch_async(dm, *pa, cur));
        }
        pre -= 1;
    } else {
        FAMG_TRY(exch_async(dm, *pa, cur));  // W-cycle revisit: the iterate was produced without a push for this plan
    }
    for (int i = 0; i < pre; ++i) {
        FAMG_TRY(dist_apply_push(dm, *L.A, EPI_SMOOTH, cur, oth, f, L.d, pa, nullptr, nullptr));
        std::swap(cur, oth);
    }
    FAMG_TRY(dist_apply_push(dm, *L.A, EPI_RESID, cur, oth, f, nullptr, &L.R->halo, nullptr, nullptr));
    const bool coarse_rep = level + 1 == dm->lrep;
    double *fc, *vc;
    if (coarse_rep) { fc = dm->fc_loc; vc = dm->g_v; }
    else { fc = dm->lv[(size_t)level + 1].b; vc = dm->lv[(size_t)level + 1].x; }
    FAMG_TRY(dist_apply_push(dm, *L.R, EPI_SPMV, oth, fc, nullptr, nullptr, nullptr, nullptr, nullptr));
    if (coarse_rep) {
        MgLevel &G = gm->lv[(size_t)dm->tail_first];
        const int64_t ldg = (G.a->nrows + 1) & ~(int64_t)1;
        FAMG_TRY(allgather_rows(dm, level + 1, fc, dm->g_f));
        for (int m = 0; m < mu; ++m) FAMG_TRY(mg_cycle(gm, (size_t)dm->tail_first, dm->g_v, ldg, dm->g_f, ldg, 1, m == 0));
    } else {
        for (int m = 0; m < mu; ++m)
            FAMG_TRY(dist_cycle_push(dm, level + 1, vc, fc, m == 0, m + 1 == mu ? &L.P->halo : nullptr, nullptr, nullptr));
    }
    FAMG_TRY(dist_apply_push(dm, *L.P, EPI_ADD, vc, cur, nullptr, nullptr, pa, nullptr, nullptr));
    for (int i = 0; i < nu; ++i) {
        const bool last = i + 1 == nu;
        FAMG_TRY(dist_apply_push(dm, *L.A, EPI_SMOOTH, cur, oth, f, L.d, last ? final_plan : pa, last ? dot_partials : nullptr,
                                 last ? dot_count : nullptr));
        std::swap(cur, oth);
    }
    if (cur != va) FAMG_FAIL(FAMG_ERR_INVALID, "internal: distributed ping-pong parity broken");
    return FAMG_OK;
}

// one visit of a distributed level; result in va (owned rows, ghost tail follows)
static famg_status dist_cycle(famg_dist_mg *dm, int level, double *va, const double *f, bool zero_guess) {
    famg_comm *cm = dm->comm;
    famg_ctx *ctx = cm->ctx;
    famg_mg *gm = dm->global;
    if (level == dm->lrep) {
        // replicated tail: gather f, run the global cycle from this level, keep the owned slice
        MgLevel &G = gm->lv[(size_t)dm->tail_first];
        const int64_t n = G.a->nrows;
        const auto &sp = dm->splits[level];
        FAMG_TRY(allgather_rows(dm, level, f, dm->g_f));
        if (!zero_guess) FAMG_TRY(allgather_rows(dm, level, va, dm->g_v));
        FAMG_TRY(mg_cycle(gm, (size_t)dm->tail_first, dm->g_v, (n + 1) & ~(int64_t)1, dm->g_f, (n + 1) & ~(int64_t)1, 1, zero_guess));
        CUDA_TRY(cudaMemcpyAsync(va, dm->g_v + sp[cm->rank], sizeof(double) * (sp[cm->rank + 1] - sp[cm->rank]), cudaMemcpyDeviceToDevice,
                                 ctx->stream));
        return FAMG_OK;
    }
    DistLevel &L = dm->lv[level];
    const int64_t nloc = L.r1 - L.r0;
    double *cur = va, *oth = L.t;
    const int nu = gm->nu, mu = gm->mu;
    auto sweep = [&]() -> famg_status {
        FAMG_TRY(dist_apply(cm, *L.A, EPI_SMOOTH, cur, oth, f, L.d, nullptr, nullptr));
        std::swap(cur, oth);
        return FAMG_OK;
    };
    int pre = nu;
    if (zero_guess) {
        if (((nu - 1) + nu) & 1) std::swap(cur, oth);
        FAMG_TRY(vec_scale_rows(ctx, L.d, f, 0, cur, 0, nloc, 1));
        pre -= 1;
    }
    for (int i = 0; i < pre; ++i) FAMG_TRY(sweep());
    FAMG_TRY(dist_apply(cm, *L.A, EPI_RESID, cur, oth, f, nullptr, nullptr, nullptr));
    const bool coarse_rep = level + 1 == dm->lrep;
    // restriction into the owned rows of level+1
    double *fc, *vc;
    if (coarse_rep) { fc = dm->fc_loc; vc = dm->g_v; }
    else { fc = dm->lv[level + 1].b; vc = dm->lv[level + 1].x; }
    FAMG_TRY(dist_apply(cm, *L.R, EPI_SPMV, oth, fc, nullptr, nullptr, nullptr, nullptr));
    if (coarse_rep) {
        MgLevel &G = gm->lv[(size_t)dm->tail_first];
        const int64_t ldg = (G.a->nrows + 1) & ~(int64_t)1;
        FAMG_TRY(allgather_rows(dm, level + 1, fc, dm->g_f));
        for (int m = 0; m < mu; ++m) FAMG_TRY(mg_cycle(gm, (size_t)dm->tail_first, dm->g_v, ldg, dm->g_f, ldg, 1, m == 0));
    } else {
        for (int m = 0; m < mu; ++m) FAMG_TRY(dist_cycle(dm, level + 1, vc, fc, m == 0));
    }
    FAMG_TRY(dist_apply(cm, *L.P, EPI_ADD, vc, cur, nullptr, nullptr, nullptr, nullptr));
    for (int i = 0; i < nu; ++i) FAMG_TRY(sweep());
    if (cur != va) FAMG_FAIL(FAMG_ERR_INVALID, "internal: distributed ping-pong parity broken");
    return FAMG_OK;
}

// Build the peer-memory exchange state for every plan (collective: all ranks call it with the same
// plan sequence).  Any failure leaves the NCCL path in place.
static famg_status p2p_setup(famg_dist_mg *d) {
    famg_comm *cm = d->comm;
    famg_ctx *ctx = cm->ctx;
    const int nr = cm->nranks, me = cm->rank;
    if (nr == 1 || nr > P2P_MAX_NB + 1) return FAMG_OK;
    std::vector<HaloPlan *> plans;
    for (auto &L : d->lv) { plans.push_back(&L.A->halo); plans.push_back(&L.R->halo); plans.push_back(&L.P->halo); }
    const int np = (int)plans.size();
    if (np == 0) return FAMG_OK;
    // arena layout: per plan [flags nr x u64 | epoch u64 | pad to 256] [recv0] [recv1] (256-byte aligned)
    auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t off = 0;
    for (HaloPlan *h : plans) {
        h->arena_flags = off; off = align(off + sizeof(unsigned long long) * (nr + 2));
        for (int b = 0; b < 2; ++b) { h->arena_recv[b] = off; off = align(off + sizeof(double) * (size_t)std::max(h->nghost, 1)); }
    }
    // collective regions: all-reduce [flags nr | epoch | pad][slots0][slots1], all-gather [flags nr | epoch | done][buf0][buf1]
    const size_t n_rep = (size_t)d->global->lv[(size_t)d->tail_first].a->nrows;
    size_t coll_off[6];
    coll_off[0] = off; off = align(off + sizeof(unsigned long long) * (nr + 2));
    coll_off[1] = off; off = align(off + sizeof(double) * (size_t)nr * P2P_AR_WIDTH);
    coll_off[2] = off; off = align(off + sizeof(double) * (size_t)nr * P2P_AR_WIDTH);
    coll_off[3] = off; off = align(off + sizeof(unsigned long long) * (nr + 2));
    coll_off[4] = off; off = align(off + sizeof(double) * (n_rep + 2));
    coll_off[5] = off; off = align(off + sizeof(double) * (n_rep + 2));
    d->arena_bytes = off;
    FAMG_TRY(dev_alloc_ipc(&d->arena, (int64_t)off));
    CUDA_TRY(cudaMemset(d->arena, 0, off));
    FAMG_TRY(dev_alloc(&d->d_p2p_err, 1));
    CUDA_TRY(cudaMemset(d->d_p2p_err, 0, sizeof(int)));
    FAMG_TRY(dev_alloc(&d->d_sig, 4));
    CUDA_TRY(cudaMemset(d->d_sig, 0, 4 * sizeof(unsigned)));
    // exchange IPC handles and layouts through NCCL all-gathers of raw bytes
    const int rec = (int)sizeof(cudaIpcMemHandle_t);
    const int lay = (int)(sizeof(long long) * ((size_t)np * (3 + nr) + 8));  // plan rows + 8-word header (collective offsets)
    unsigned char *dbuf = nullptr;
    FAMG_TRY(dev_alloc(&dbuf, (int64_t)(rec + lay) * (nr + 1)));
    std::vector<unsigned char> mine((size_t)(rec + lay));
    cudaIpcMemHandle_t hnd;
    CUDA_TRY(cudaIpcGetMemHandle(&hnd, d->arena));
    memcpy(mine.data(), &hnd, rec);
    long long *tab = reinterpret_cast<long long *>(mine.data() + rec) + 8;
    for (int w = 0; w < 6; ++w) tab[w - 8] = (long long)coll_off[w];
    for (int i = 0; i < np; ++i) {
        long long *row = tab + (size_t)i * (3 + nr);
        row[0] = (long long)plans[i]->arena_flags; row[1] = (long long)plans[i]->arena_recv[0]; row[2] = (long long)plans[i]->arena_recv[1];
        for (int p = 0; p < nr; ++p) row[3 + p] = plans[i]->recv_off[p];
    }
    unsigned char *dsend = dbuf + (size_t)(rec + lay) * nr;
    CUDA_TRY(cudaMemcpy(dsend, mine.data(), mine.size(), cudaMemcpyHostToDevice));
    NCCL_TRY(g_nccl.AllGather(dsend, dbuf, (size_t)(rec + lay), ncclInt8, cm->comm, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    std::vector<unsigned char> all((size_t)(rec + lay) * nr);
    CUDA_TRY(cudaMemcpy(all.data(), dbuf, all.size(), cudaMemcpyDeviceToHost));
    cudaFree(dbuf);
    d->peer_arena.assign(nr, nullptr);
    for (int p = 0; p < nr; ++p) {
        if (p == me) { d->peer_arena[p] = d->arena; continue; }
        cudaIpcMemHandle_t ph;
        memcpy(&ph, all.data() + (size_t)(rec + lay) * p, rec);
        cudaError_t e = cudaIpcOpenMemHandle(&d->peer_arena[p], ph, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { cudaGetLastError(); FAMG_FAIL(FAMG_ERR_COMM, "cudaIpcOpenMemHandle(rank %d) failed: %s", p, cudaGetErrorString(e)); }
    }
    for (int i = 0; i < np; ++i) {
        HaloPlan *h = plans[i];
        P2PPlanDev &v = h->dev;
        v = P2PPlanDev{};
        v.nghost = h->nghost;
        v.epoch = reinterpret_cast<unsigned long long *>(d->arena + h->arena_flags) + nr;
        v.done = reinterpret_cast<unsigned int *>(reinterpret_cast<unsigned long long *>(d->arena + h->arena_flags) + nr + 1);
        v.total_send = h->total_send;
        v.lrecv[0] = reinterpret_cast<const double *>(d->arena + h->arena_recv[0]);
        v.lrecv[1] = reinterpret_cast<const double *>(d->arena + h->arena_recv[1]);
        v.err = d->d_p2p_err;
        for (int p = 0; p < nr; ++p) {
            if (p == me || (h->send_cnt[p] == 0 && h->recv_cnt[p] == 0)) continue;
            const long long *prow = reinterpret_cast<const long long *>(all.data() + (size_t)(rec + lay) * p + rec) + 8 + (size_t)i * (3 + nr);
            unsigned char *pa = static_cast<unsigned char *>(d->peer_arena[p]);
            const int nb = v.nnb++;
            v.rdst[0][nb] = reinterpret_cast<double *>(pa + prow[1]) + prow[3 + me];
            v.rdst[1][nb] = reinterpret_cast<double *>(pa + prow[2]) + prow[3 + me];
            v.rflag[nb] = reinterpret_cast<unsigned long long *>(pa + prow[0]) + me;
            v.lflag[nb] = reinterpret_cast<const unsigned long long *>(d->arena + h->arena_flags) + p;
            v.soff[nb] = h->send_off[p]; v.scnt[nb] = h->send_cnt[p];
        }
        h->p2p = h->any && v.nnb > 0;
        // fused push: device copy of the plan and the (neighbour, slot) of every boundary row
        if (h->p2p) {
            const int nb_rows = h->push_lo + (h->nloc - h->push_hi);
            std::vector<int> map((size_t)std::max(nb_rows, 1), -1);
            bool fusable = (int)h->send_idx_host.size() == h->total_send && h->total_send > 0;
            for (int nb = 0; nb < v.nnb && fusable; ++nb) {
                if (v.scnt[nb] >= (1 << PUSH_SLOT_BITS)) { fusable = false; break; }
                for (int j = 0; j < v.scnt[nb]; ++j) {
                    const int i = h->send_idx_host[(size_t)(v.soff[nb] + j)];
                    const int mi = i < h->push_lo ? i : (i >= h->push_hi ? h->push_lo + (i - h->push_hi) : -1);
                    if (mi < 0 || mi >= nb_rows || map[(size_t)mi] != -1) { fusable = false; break; }  // a row two peers need: pack kernel instead
                    map[(size_t)mi] = (nb << PUSH_SLOT_BITS) | j;
                }
            }
            for (int nb = 0; nb < v.nnb && fusable; ++nb) fusable = v.scnt[nb] > 0;  // somebody has to publish towards every neighbour
            // Measured on 8 B200 (profiles/r2_timeline_*): when most rows of a slab are boundary rows (thin slabs of the coarse
            // levels) the per-warp system fences and counter updates cost the producer more than the pack kernel they replace.
            if (fusable && (int64_t)nb_rows * 100 > (int64_t)h->nloc * d->fuse_max_boundary_pct) fusable = false;
            if (fusable) {
                FAMG_TRY(dev_alloc(&h->d_dev, 1));
                FAMG_TRY(dev_alloc(&h->d_push_map, (int64_t)map.size()));
                CUDA_TRY(cudaMemcpy(h->d_dev, &v, sizeof(v), cudaMemcpyHostToDevice));
                CUDA_TRY(cudaMemcpy(h->d_push_map, map.data(), sizeof(int) * map.size(), cudaMemcpyHostToDevice));
            }
            h->fusable = fusable;
        }
    }
    {   // collective descriptors
        P2PCollDev &c = d->coll;
        c = P2PCollDev{};
        c.nranks = nr; c.rank = me; c.err = d->d_p2p_err;
        for (int p = 0; p < nr; ++p) {
            const long long *hdr = reinterpret_cast<const long long *>(all.data() + (size_t)(rec + lay) * p + rec);
            unsigned char *pa = static_cast<unsigned char *>(d->peer_arena[p]);
            c.ar_rslots[0][p] = reinterpret_cast<double *>(pa + hdr[1]);
            c.ar_rslots[1][p] = reinterpret_cast<double *>(pa + hdr[2]);
            c.ar_rflag[p] = reinterpret_cast<unsigned long long *>(pa + hdr[0]) + me;
            c.ag_rbuf[0][p] = reinterpret_cast<double *>(pa + hdr[4]);
            c.ag_rbuf[1][p] = reinterpret_cast<double *>(pa + hdr[5]);
            c.ag_rflag[p] = reinterpret_cast<unsigned long long *>(pa + hdr[3]) + me;
        }
        c.ar_lflag = reinterpret_cast<const unsigned long long *>(d->arena + coll_off[0]);
        c.ar_epoch = reinterpret_cast<unsigned long long *>(d->arena + coll_off[0]) + nr;
        c.ar_lslots[0] = reinterpret_cast<const double *>(d->arena + coll_off[1]);
        c.ar_lslots[1] = reinterpret_cast<const double *>(d->arena + coll_off[2]);
        c.ag_lflag = reinterpret_cast<const unsigned long long *>(d->arena + coll_off[3]);
        c.ag_epoch = reinterpret_cast<unsigned long long *>(d->arena + coll_off[3]) + nr;
        c.ag_done = reinterpret_cast<unsigned int *>(reinterpret_cast<unsigned long long *>(d->arena + coll_off[3]) + nr + 1);
        c.ag_lbuf[0] = reinterpret_cast<const double *>(d->arena + coll_off[4]);
        c.ag_lbuf[1] = reinterpret_cast<const double *>(d->arena + coll_off[5]);
        d->p2p_coll = getenv("FAMG_P2P_COLL") ? atoi(getenv("FAMG_P2P_COLL")) != 0 : true;
    }
    d->p2p = true;
    return FAMG_OK;
}

static void dist_free(famg_dist_mg *d) {
    for (size_t p = 0; p < d->peer_arena.size(); ++p)
        if (d->peer_arena[p] && (int)p != d->comm->rank) cudaIpcCloseMemHandle(d->peer_arena[p]);
    cudaFree(d->arena); cudaFree(d->d_p2p_err); cudaFree(d->d_sig);
    for (auto &l : d->lv) { cudaFree(l.d); cudaFree(l.x); cudaFree(l.b); cudaFree(l.t); }
    for (famg_dmat *m : d->keep) dmat_release(m);
    if (d->owns_tail && d->global) famg_mg_destroy(d->global);
    for (auto &g : d->graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
    cudaFree(d->g_f); cudaFree(d->g_v); cudaFree(d->fc_loc); cudaFree(d->pcg);
    delete d;
}

}  // namespace famg

using namespace famg;

extern "C" {

famg_status famg_comm_unique_id(void *id_bytes) {
    if (!id_bytes) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    FAMG_TRY(nccl_load());
    static_assert(sizeof(ncclUniqueId) == FAMG_UNIQUE_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(id_bytes, &id, sizeof(id));
    return FAMG_OK;
}

famg_status famg_comm_create(famg_ctx *ctx, int nranks, int rank, const void *id_bytes, famg_comm **out) {
    if (!ctx || !out || nranks < 1 || rank < 0 || rank >= nranks) FAMG_FAIL(FAMG_ERR_INVALID, "bad argument");
    *out = nullptr;
    CUDA_TRY(cudaSetDevice(ctx->device));
    famg_comm *c = new famg_comm();
    c->ctx = ctx; c->nranks = nranks; c->rank = rank;
    if (nranks > 1) {
        if (!id_bytes) { delete c; FAMG_FAIL(FAMG_ERR_INVALID, "null unique id"); }
        famg_status st = nccl_load();
        if (st != FAMG_OK) { delete c; return st; }
        ncclUniqueId id;
        memcpy(&id, id_bytes, sizeof(id));
        ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
        if (r != ncclSuccess) { delete c; FAMG_FAIL(FAMG_ERR_COMM, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)); }
    }
    cudaEventCreateWithFlags(&c->ev_packed, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_packed_e, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_halo_e, cudaEventDisableTiming);
    *out = c;
    return FAMG_OK;
}

famg_status famg_comm_destroy(famg_comm *c) {
    if (!c) return FAMG_OK;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream); cudaStreamSynchronize(c->ctx->comm_stream);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    cudaEventDestroy(c->ev_packed); cudaEventDestroy(c->ev_halo);
    if (c->ev_packed_e) cudaEventDestroy(c->ev_packed_e);
    if (c->ev_halo_e) cudaEventDestroy(c->ev_halo_e);
    delete c;
    return FAMG_OK;
}

famg_status famg_comm_allreduce_sum(famg_comm *c, double *vals, int n) {
    if (!c || !vals || n < 0 || n > 32) FAMG_FAIL(FAMG_ERR_INVALID, "bad argument (n <= 32)");
    if (c->nlocal > 1) FAMG_FAIL(FAMG_ERR_UNSUPPORTED, "virtual communicators have no collectives outside the hierarchy setup");
    famg_ctx *ctx = c->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaMemcpyAsync(ctx->d_scalars + 32, vals, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    FAMG_TRY(allreduce_slots(c, 32, n));
    CUDA_TRY(cudaMemcpyAsync(vals, ctx->d_scalars + 32, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return FAMG_OK;
}

// Work vectors, transition buffers and the peer-memory state once d->lv[*].{A,R,P}, d->splits, d->global and
// d->tail_first are in place (both constructors end here).
static famg_status dist_mg_finish(famg_dist_mg *d, int diag_kind, double omega, const famg_mg *diag_source) {
    famg_comm *c = d->comm;
    famg_ctx *ctx = c->ctx;
    const int lrep = d->lrep;
    if (const char *v = getenv("FAMG_DIST_GRAPH")) d->use_graph = atoi(v) != 0;
    if (const char *v = getenv("FAMG_OVERLAP")) { d->overlap_mode = std::min(std::max(atoi(v), 0), 3); d->overlap = d->overlap_mode != 0; }
    if (const char *v = getenv("FAMG_OVERLAP_MIN_ROWS")) d->split_min_rows = std::max(atoi(v), 1);
    if (const char *v = getenv("FAMG_RESERVE_CTAS")) d->reserve_ctas = std::max(atoi(v), 0);
    if (const char *v = getenv("FAMG_FUSE_MAX_BOUNDARY_PCT")) d->fuse_max_boundary_pct = std::max(atoi(v), 0);
    FAMG_TRY(mg_ensure_workspace(d->global, 1));
    famg_status st = FAMG_OK;
    for (int l = 0; l < lrep && st == FAMG_OK; ++l) {
        DistLevel &L = d->lv[(size_t)l];
        L.r0 = d->splits[(size_t)l][(size_t)c->rank]; L.r1 = d->splits[(size_t)l][(size_t)c->rank + 1];
        const int64_t nloc = L.r1 - L.r0;
        st = dev_alloc(&L.d, nloc);
        if (st != FAMG_OK) break;
        if (diag_source) {  // slice of the replicated level smoother
            if (nloc) cudaMemcpyAsync(L.d, diag_source->lv[(size_t)l].s->d + L.r0, sizeof(double) * nloc, cudaMemcpyDeviceToDevice, ctx->stream);
        } else {
            st = diag_from_rows(L.A->local, diag_kind, omega, L.d);
        }
    }
    // vector tails: a level-l vector is consumed by A_l and R_l (level l columns) and by P_{l-1}
    for (int l = 0; l < lrep && st == FAMG_OK; ++l) {
        DistLevel &L = d->lv[(size_t)l];
        int64_t ghost = std::max(L.A->halo.nghost, L.R->halo.nghost);
        if (l > 0) ghost = std::max<int64_t>(ghost, d->lv[(size_t)l - 1].P->halo.nghost);
        L.ld = (L.r1 - L.r0) + ghost + 2;
        st = dev_alloc(&L.x, L.ld);
        if (st == FAMG_OK) st = dev_alloc(&L.b, L.ld);
        if (st == FAMG_OK) st = dev_alloc(&L.t, L.ld);
        if (st == FAMG_OK) {
            cudaMemsetAsync(L.x, 0, sizeof(double) * L.ld, ctx->stream);
            cudaMemsetAsync(L.b, 0, sizeof(double) * L.ld, ctx->stream);
            cudaMemsetAsync(L.t, 0, sizeof(double) * L.ld, ctx->stream);
        }
    }
    if (st == FAMG_OK) {
        const int64_t n = d->global->lv[(size_t)d->tail_first].a->nrows;
        st = dev_alloc(&d->g_f, n + 2);
        if (st == FAMG_OK) st = dev_alloc(&d->g_v, n + 2);
        if (st == FAMG_OK) st = dev_alloc(&d->fc_loc, d->splits[(size_t)lrep][(size_t)c->rank + 1] - d->splits[(size_t)lrep][(size_t)c->rank] + 2);
    }
    if (st == FAMG_OK) {
        // PCG vectors r, z, q, x-scratch, and p with the ghost tail of A_0
        const int64_t nloc = d->splits[0][(size_t)c->rank + 1] - d->splits[0][(size_t)c->rank];
        const int64_t ghost = lrep > 0 ? d->lv[0].ld : nloc + 2;
        d->pcg_ld = (std::max(nloc + 2, ghost) + 1) & ~(int64_t)1;
        st = dev_alloc(&d->pcg, 6 * d->pcg_ld);
        if (st == FAMG_OK) cudaMemsetAsync(d->pcg, 0, sizeof(double) * 6 * d->pcg_ld, ctx->stream);
    }
    cudaStreamSynchronize(ctx->stream);
    FAMG_TRY(st);
    {   // peer-memory halo exchange unless FAMG_HALO=nccl; every rank must reach the same decision
        const char *mode = getenv("FAMG_HALO");
        if (!(mode && !strcmp(mode, "nccl"))) {
            famg_status ps = p2p_setup(d);
            double ok = (ps == FAMG_OK && d->p2p) ? 1.0 : 0.0, all_ok = ok;
            if (c->nranks > 1) {
                double v[1] = {ok};
                famg_status as = famg_comm_allreduce_sum(c, v, 1);
                all_ok = (as == FAMG_OK && v[0] == (double)c->nranks) ? 1.0 : 0.0;
            }
            if (all_ok != 1.0) {  // somebody could not map a peer: everyone stays on NCCL send/recv
                for (auto &L : d->lv) { L.A->halo.p2p = false; L.R->halo.p2p = false; L.P->halo.p2p = false; }
                d->p2p = false; d->p2p_coll = false;
            }
        }
    }
    return FAMG_OK;
}

// row slab [rs[rank], rs[rank+1]) of a replicated matrix as a one-part distributed matrix
static famg_status dmat_from_replicated(famg_comm *c, const famg_csr *m, const std::vector<int64_t> &rs, const std::vector<int64_t> &cs,
                                        bool replicated_cols, famg_dmat **out) {
    famg_csr *slab = nullptr;
    FAMG_TRY(famg_csr_row_slab(m, rs[(size_t)c->rank], rs[(size_t)c->rank + 1], &slab));
    famg_dmat *dm = new famg_dmat();
    dm->comm = c; dm->nrows = m->nrows; dm->ncols = m->ncols; dm->rsplit = rs; dm->csplit = cs;
    dm->part.resize(1);
    dm->part[0].local = slab;
    famg_status st = dmat_finalize(dm, replicated_cols);
    if (st != FAMG_OK) { dmat_release(dm); return st; }
    *out = dm;
    return FAMG_OK;
}

famg_status famg_dist_mg_create(famg_comm *c, famg_mg *gm, const int64_t *const *row_splits, int64_t replicate_below,
                                famg_dist_mg **out) {
    if (!c || !gm || !row_splits || !out) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    *out = nullptr;
    famg_ctx *ctx = c->ctx;
    if (gm->ctx != ctx) FAMG_FAIL(FAMG_ERR_INVALID, "multigrid and communicator live on different contexts");
    if (c->nlocal != 1) FAMG_FAIL(FAMG_ERR_UNSUPPORTED, "a single-process virtual communicator can build hierarchies but not run the distributed cycle");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int nl = (int)gm->lv.size(), nr = c->nranks;
    famg_dist_mg *d = new famg_dist_mg();
    d->comm = c; d->global = gm;
    d->splits.resize((size_t)nl);
    for (int l = 0; l < nl; ++l) {
        d->splits[(size_t)l].assign(row_splits[l], row_splits[l] + nr + 1);
        const auto &sp = d->splits[(size_t)l];
        bool ok = sp[0] == 0 && sp[(size_t)nr] == gm->lv[(size_t)l].a->nrows;
        for (int p = 0; p < nr; ++p) ok = ok && sp[(size_t)p] <= sp[(size_t)p + 1];
        if (!ok) { dist_free(d); FAMG_FAIL(FAMG_ERR_INVALID, "row_splits of level %d do not partition its rows", l); }
    }
    // first replicated level: too few rows per rank, a non-Diag smoother, or the coarsest level
    int lrep = nl - 1;
    for (int l = 0; l < nl - 1; ++l)
        if (gm->lv[(size_t)l].a->nrows / nr < replicate_below || gm->lv[(size_t)l].s->kind != SM_DIAG) { lrep = l; break; }
    d->lrep = lrep; d->tail_first = lrep;
    d->lv.resize((size_t)lrep);
    famg_status st = FAMG_OK;
    for (int l = 0; l < lrep && st == FAMG_OK; ++l) {
        DistLevel &L = d->lv[(size_t)l];
        MgLevel &G = gm->lv[(size_t)l], &C = gm->lv[(size_t)l + 1];
        const bool coarse_rep = l + 1 == lrep;
        famg_dmat *A = nullptr, *R = nullptr, *P = nullptr;
        st = dmat_from_replicated(c, G.a, d->splits[(size_t)l], d->splits[(size_t)l], false, &A);
        if (st == FAMG_OK) { d->keep.push_back(A); L.A = &A->part[0]; st = dmat_from_replicated(c, C.r, d->splits[(size_t)l + 1], d->splits[(size_t)l], false, &R); }
        if (st == FAMG_OK) { d->keep.push_back(R); L.R = &R->part[0]; st = dmat_from_replicated(c, C.p, d->splits[(size_t)l], d->splits[(size_t)l + 1], coarse_rep, &P); }
        if (st == FAMG_OK) { d->keep.push_back(P); L.P = &P->part[0]; }
    }
    if (st == FAMG_OK) st = dist_mg_finish(d, FAMG_DIAG_L1, 0.0, gm);
    if (st != FAMG_OK) { dist_free(d); return st; }
    *out = d;
    return FAMG_OK;
}

// Level-wise construction from distributed operators (famg_dist_coarsen): a[l] (finalized), r[l] (finalized over
// a[l]'s split), p[l] (finalized over a[l+1]'s split, or with replicated columns for the last one), l < nlevels;
// `tail` is the replicated Multigrid of the remaining levels (its level 0 = the gathered a[nlevels]); it is
// borrowed and must outlive the result.  Smoother of the distributed levels: FAMG_DIAG_L1 | FAMG_DIAG_JACOBI.
famg_status famg_dist_mg_create_levels(famg_comm *c, int nlevels, famg_dmat *const *a, famg_dmat *const *r, famg_dmat *const *p,
                                       int diag_kind, double omega, famg_mg *tail, famg_dist_mg **out) {
    if (!c || !tail || !out || nlevels < 0 || (nlevels > 0 && (!a || !r || !p))) FAMG_FAIL(FAMG_ERR_INVALID, "bad argument");
    *out = nullptr;
    famg_ctx *ctx = c->ctx;
    if (tail->ctx != ctx) FAMG_FAIL(FAMG_ERR_INVALID, "multigrid and communicator live on different contexts");
    if (c->nlocal != 1) FAMG_FAIL(FAMG_ERR_UNSUPPORTED, "a single-process virtual communicator can build hierarchies but not run the distributed cycle");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int nr = c->nranks;
    for (int l = 0; l < nlevels; ++l) {
        if (!a[l] || !r[l] || !p[l] || a[l]->comm != c || r[l]->comm != c || p[l]->comm != c) FAMG_FAIL(FAMG_ERR_INVALID, "level %d: null operator or foreign communicator", l);
        if (!a[l]->finalized || a[l]->replicated_cols || !r[l]->finalized || r[l]->replicated_cols || !p[l]->finalized)
            FAMG_FAIL(FAMG_ERR_INVALID, "level %d: operators must be finalized (famg_dmat_finalize)", l);
        const bool last = l + 1 == nlevels;
        if (p[l]->replicated_cols != last) FAMG_FAIL(FAMG_ERR_INVALID, "level %d: only the last prolongator has replicated columns", l);
        const int64_t nc = last ? tail->lv[0].a->nrows : a[l + 1]->nrows;
        if (a[l]->nrows != a[l]->ncols || r[l]->ncols != a[l]->nrows || p[l]->nrows != a[l]->nrows || r[l]->nrows != nc || p[l]->ncols != nc ||
            a[l]->rsplit != a[l]->csplit || r[l]->csplit != a[l]->rsplit || p[l]->rsplit != a[l]->rsplit || (!last && (r[l]->rsplit != a[l + 1]->rsplit || p[l]->csplit != a[l + 1]->rsplit)))
            FAMG_FAIL(FAMG_ERR_INVALID, "level %d: shape / split mismatch between A, R and P (hierarchy.rs:258-264)", l);
    }
    famg_dist_mg *d = new famg_dist_mg();
    d->comm = c; d->global = tail; d->tail_first = 0; d->lrep = nlevels;
    d->splits.resize((size_t)nlevels + 1);
    d->lv.resize((size_t)nlevels);
    for (int l = 0; l < nlevels; ++l) {
        d->splits[(size_t)l] = a[l]->rsplit;
        for (famg_dmat *m : {a[l], r[l], p[l]}) { m->refs.fetch_add(1); d->keep.push_back(m); }
        d->lv[(size_t)l].A = &a[l]->part[0]; d->lv[(size_t)l].R = &r[l]->part[0]; d->lv[(size_t)l].P = &p[l]->part[0];
    }
    if (nlevels > 0) d->splits[(size_t)nlevels] = r[nlevels - 1]->rsplit;
    else { d->splits[0].assign((size_t)nr + 1, tail->lv[0].a->nrows); d->splits[0][0] = 0; }
    famg_status st = dist_mg_finish(d, diag_kind, omega, nullptr);
    if (st != FAMG_OK) { dist_free(d); return st; }
    *out = d;
    return FAMG_OK;
}

famg_status famg_dist_mg_destroy(famg_dist_mg *d) {
    if (!d) return FAMG_OK;
    cudaSetDevice(d->comm->ctx->device);
    cudaStreamSynchronize(d->comm->ctx->stream); cudaStreamSynchronize(d->comm->ctx->comm_stream);
    dist_free(d);
    return FAMG_OK;
}

static famg_status dist_check_local(const famg_dist_mg *d, const famg_vec *v) {
    const int64_t nloc = d->splits[0][d->comm->rank + 1] - d->splits[0][d->comm->rank];
    if (!v || v->nrows != nloc || v->ncols != 1) FAMG_FAIL(FAMG_ERR_INVALID, "distributed vectors must hold this rank's %lld rows (1 column)", (long long)nloc);
    return FAMG_OK;
}

// out_local = B rhs_local: one distributed mu-cycle from a zero guess. (buffers: pcg slot 4/5)
// With dot_partials (peer-memory overlap mode) the last sweep also leaves *dot_count partial sums of rhs . out there.
static famg_status dist_precond(famg_dist_mg *d, double *out_ext, const double *rhs, double *dot_partials = nullptr, int *dot_count = nullptr) {
    famg_ctx *ctx = d->comm->ctx;
    if (dot_count) *dot_count = 0;
    const bool push = d->overlap && d->p2p && d->lrep > 0;
    if (!push) dot_partials = nullptr;
    auto run = [&](int *nd) -> famg_status {
        if (push) return dist_cycle_push(d, 0, out_ext, rhs, true, nullptr, dot_partials, dot_partials ? nd : nullptr);
        return dist_cycle(d, 0, out_ext, rhs, true);
    };
    if (!d->use_graph) return run(dot_count);
    auto key = std::make_tuple((const void *)out_ext, (const void *)rhs, (const void *)dot_partials);
    auto it = d->graphs.find(key);
    if (it == d->graphs.end()) {
        const int64_t before = ctx->launches.load();
        cudaError_t e = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
        famg_status st = FAMG_ERR_CUDA;
        cudaGraph_t graph = nullptr;
        int nd = 0;
        if (e == cudaSuccess) {
            st = run(&nd);
            e = cudaStreamEndCapture(ctx->stream, &graph);
        }
        d->pending = false;
        GraphEntry ent;
        if (st == FAMG_OK && e == cudaSuccess && graph) e = cudaGraphInstantiate(&ent.exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (st != FAMG_OK || e != cudaSuccess || !ent.exec) {
            // capture is an optimisation: fall back to eager launches for the rest of this object's life
            cudaGetLastError();
            d->use_graph = false;
            ctx->launches.store(before);
            return run(dot_count);
        }
        ent.launches = ctx->launches.load() - before;
        ent.dot_count = nd;
        ctx->launches.store(before);
        it = d->graphs.emplace(key, ent).first;
    }
    CUDA_TRY(cudaGraphLaunch(it->second.exec, ctx->stream));
    count_launch(ctx, (int)it->second.launches);
    if (dot_count) *dot_count = it->second.dot_count;
    return FAMG_OK;
}

famg_status famg_dist_mg_apply_dev(famg_dist_mg *d, famg_vec *out_local, const famg_vec *rhs_local) {
    if (!d) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    FAMG_TRY(dist_check_local(d, out_local));
    FAMG_TRY(dist_check_local(d, rhs_local));
    famg_ctx *ctx = d->comm->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    double *z = d->pcg + 4 * d->pcg_ld;
    FAMG_TRY(dist_precond(d, z, rhs_local->p));
    return vec_copy(ctx, out_local->p, out_local->ld, z, d->pcg_ld, out_local->nrows, 1);
}

famg_status famg_dist_spmv_dev(famg_dist_mg *d, famg_vec *y_local, const famg_vec *x_local) {
    if (!d) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    FAMG_TRY(dist_check_local(d, y_local));
    FAMG_TRY(dist_check_local(d, x_local));
    famg_ctx *ctx = d->comm->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (d->lrep == 0) FAMG_FAIL(FAMG_ERR_UNSUPPORTED, "level 0 is replicated on this configuration");
    double *xe = d->pcg + 5 * d->pcg_ld;
    FAMG_TRY(vec_copy(ctx, xe, d->pcg_ld, x_local->p, x_local->ld, x_local->nrows, 1));
    return dist_apply(d->comm, *d->lv[0].A, EPI_SPMV, xe, y_local->p, nullptr, nullptr, nullptr, nullptr);
}

famg_status famg_dist_pcg_solve_dev(famg_dist_mg *d, famg_vec *x, const famg_vec *b, double rel_tol, double abs_tol,
                                    int64_t max_iters, int zero_guess, famg_cg_info *info) {
    if (!d || !info) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    FAMG_TRY(dist_check_local(d, x));
    FAMG_TRY(dist_check_local(d, b));
    famg_comm *cm = d->comm;
    famg_ctx *ctx = cm->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (d->lrep == 0) FAMG_FAIL(FAMG_ERR_UNSUPPORTED, "problem too small to partition: use famg_pcg_solve on one GPU");
    const int64_t n = x->nrows, ld = d->pcg_ld;
    info->iter_count = 0; info->abs_residual = 0; info->rel_residual = 0;
    double *r = d->pcg, *p = d->pcg + ld, *z = d->pcg + 2 * ld, *q = d->pcg + 3 * ld, *xe = d->pcg + 5 * ld;
    DistOp &A = *d->lv[0].A;
    enum { S_BB = 0, S_RR = 1, S_PTQ = 2, S_RTZ_A = 3, S_RTZ_B = 4 };
    double h[8];
    FAMG_TRY(vec_dot(ctx, b->p, b->p, n, S_BB));
    FAMG_TRY(dist_allreduce(d, S_BB, 1));
    FAMG_TRY(read_scalars(ctx, S_BB, 1, h));
    const double b_norm = sqrt(h[0]);
    if (b_norm == 0.0) return famg_vec_fill(x, 0.0);
    const double thr = std::max(abs_tol, rel_tol * b_norm);
    if (zero_guess) {
        FAMG_TRY(famg_vec_fill(x, 0.0));
        FAMG_TRY(vec_copy(ctx, r, ld, b->p, b->ld, n, 1));
    } else {
        FAMG_TRY(vec_copy(ctx, xe, ld, x->p, x->ld, n, 1));
        FAMG_TRY(dist_apply(cm, A, EPI_RESID, xe, r, b->p, nullptr, nullptr, nullptr));
    }
    FAMG_TRY(vec_dot(ctx, r, r, n, S_RR));
    FAMG_TRY(dist_allreduce(d, S_RR, 1));
    FAMG_TRY(read_scalars(ctx, S_RR, 1, h));
    double rn = sqrt(h[0]);
    bool converged = rn < thr;
    int slot_rtz = S_RTZ_A, slot_rtz_new = S_RTZ_B;
    const bool push = d->overlap && d->p2p;
    // p = z (+ beta p), with the exchange of p started as soon as the entries the neighbours need are written
    auto update_p = [&](bool first, int slot_new, int slot_old) -> famg_status {
        const HaloPlan &h = A.halo;
        const int lo = h.push_lo, hi = h.push_hi;
        auto piece = [&](int64_t off, int64_t cnt) -> famg_status {
            if (cnt <= 0) return FAMG_OK;
            if (first) return vec_copy(ctx, p + off, ld, z + off, ld, cnt, 1);
            return pcg_update_p(ctx, p + off, z + off, cnt, slot_new, slot_old);
        };
        if (!push) return piece(0, n);
        if (hi - lo < d->split_min_rows) { FAMG_TRY(piece(0, n)); return exch_async(d, h, p); }
        FAMG_TRY(piece(0, lo));
        FAMG_TRY(piece(hi, n - hi));
        FAMG_TRY(exch_async(d, h, p));
        return piece(lo, hi - lo);
    };
    auto rtz = [&](int nd, int slot) -> famg_status {
        if (nd > 0) return dist_reduce_allreduce(d, nd, slot);
        FAMG_TRY(vec_dot(ctx, r, z, n, slot));
        return dist_allreduce(d, slot, 1);
    };
    if (!converged) {
        FAMG_TRY(ensure_partials(ctx, n + 8));
        int nd = 0;
        FAMG_TRY(dist_precond(d, z, r, ctx->d_partials, &nd));
        FAMG_TRY(rtz(nd, slot_rtz));
        FAMG_TRY(update_p(true, 0, 0));
        for (int64_t it = 0; it < max_iters; ++it) {
            int np = 0;
            if (push) FAMG_TRY(dist_apply_push(d, A, EPI_SPMV, p, q, nullptr, nullptr, nullptr, ctx->d_partials, &np));
            else FAMG_TRY(dist_apply(cm, A, EPI_SPMV, p, q, nullptr, nullptr, ctx->d_partials, &np));
            FAMG_TRY(dist_reduce_allreduce(d, np, S_PTQ));
            FAMG_TRY(pcg_update_xr_partials(ctx, x->p, r, p, q, n, slot_rtz, S_PTQ, &np));
            FAMG_TRY(dist_reduce_allreduce(d, np, S_RR));
            FAMG_TRY(read_scalars(ctx, S_RR, 4, h));
            const double ptq = h[S_PTQ - S_RR], rtzv = h[slot_rtz - S_RR];
            if (!(ptq > 0.0) || !(rtzv > 0.0))
                FAMG_FAIL(FAMG_ERR_NOT_SPD, "pcg: operator or preconditioner is not positive definite (p.Ap=%g, r.z=%g)", ptq, rtzv);
            rn = sqrt(h[0]);
            info->iter_count = it + 1;
            if (rn < thr) { converged = true; break; }
            FAMG_TRY(dist_precond(d, z, r, ctx->d_partials, &nd));
            FAMG_TRY(rtz(nd, slot_rtz_new));
            FAMG_TRY(update_p(false, slot_rtz_new, slot_rtz));
            std::swap(slot_rtz, slot_rtz_new);
        }
        FAMG_TRY(exch_wait(d));
    }
    info->abs_residual = rn; info->rel_residual = rn / b_norm;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (d->p2p) {
        int perr = 0;
        CUDA_TRY(cudaMemcpy(&perr, d->d_p2p_err, sizeof(int), cudaMemcpyDeviceToHost));
        if (perr) FAMG_FAIL(FAMG_ERR_COMM, "peer-memory halo exchange timed out waiting for a neighbour rank");
    }
    if (!converged) FAMG_FAIL(FAMG_ERR_NO_CONVERGENCE, "pcg: no convergence in %lld iterations (abs %.3e, rel %.3e)", (long long)max_iters, rn, rn / b_norm);
    return FAMG_OK;
}

famg_status famg_dist_pcg_solve(famg_dist_mg *d, double *x_local, const double *b_local, double rel_tol, double abs_tol,
                                int64_t max_iters, int zero_guess, famg_cg_info *info) {
    if (!d || !x_local || !b_local || !info) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    famg_ctx *ctx = d->comm->ctx;
    const int64_t n = d->splits[0][d->comm->rank + 1] - d->splits[0][d->comm->rank];
    famg_vec *dx = nullptr, *db = nullptr;
    FAMG_TRY(famg_vec_create(ctx, n, 1, &dx));
    famg_status st = famg_vec_create(ctx, n, 1, &db);
    if (st == FAMG_OK) st = famg_vec_upload(db, b_local, n);
    if (st == FAMG_OK && !zero_guess) st = famg_vec_upload(dx, x_local, n);
    famg_status solve = FAMG_OK;
    if (st == FAMG_OK) {
        solve = famg_dist_pcg_solve_dev(d, dx, db, rel_tol, abs_tol, max_iters, zero_guess, info);
        if (solve == FAMG_OK || solve == FAMG_ERR_NO_CONVERGENCE) {
            std::string keep = get_error();
            st = famg_vec_download(dx, x_local, n);
            if (solve != FAMG_OK) set_error("%s", keep.c_str());
        }
    }
    famg_vec_destroy(dx); famg_vec_destroy(db);
    return st != FAMG_OK ? st : solve;
}

}  // extern "C"
