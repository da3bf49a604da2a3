// spmv.cu -- the CSR mat-apply family: y = A x, r = b - A x, fused diagonal smoother sweep,
// y += A x, and the stationary-iteration step; one or many right-hand sides.
//
// Replaces ParSpmmOp::apply / BlockRow::spmm (par_spmm.rs:98-133), the serial CSR LinOp used for
// R and P (multigrid.rs:157-159), and the unfused smooth()/residual passes (multigrid.rs:341-350,
// 407-424).
//
// Kernel design (HBM-bound, no tensor cores):
//   * one CTA of 256 threads owns ROWS = 256/TPR consecutive rows; TPR (threads per row: 1, 2, 4,
//     ..., 32) is picked per matrix from its row-length statistics (csr_finalize_plan) so that a
//     CTA covers ~2k non-zeros: scalar-per-row for 7-point rows, 4 lanes per row for 27-point
//     rows, a full warp per row for >128 non-zeros per row.
//   * phase 1: the CTA's contiguous slice of val[] / col[] is streamed global -> shared with
//     fully coalesced 128-bit loads (ld.global.nc.L1::no_allocate), independent of row shape.
//   * phase 2: each row group walks its row out of shared memory and gathers x through L1/L2.
//     With TPR = 1 adjacent threads read adjacent rows, so for stencil matrices every gather
//     instruction touches 1-2 cache lines; the adds happen in ascending column order, i.e. the
//     exact summation order of the reference (bit-identical results when TPR == 1).
//   * the epilogue fuses the vector work of the caller (residual, smoother update,
//     prolongation-add, p.Ap partial) so each sweep reads A once and every vector once.
//   * a slice that does not fit the staging buffer (very irregular rows) is processed straight
//     from global memory by the same row groups -- slower, still correct.
// Compiled with -fmad=false: products and sums are rounded separately, like the reference's
// scalar `dst += a*b` loops.
#include "common.cuh"
#include "p2p_dev.cuh"

namespace famg {

constexpr int SPMV_THREADS = 256;
constexpr int SPMV_CAP = 2560;  // staged non-zeros per CTA (30 KB: 20 KB values + 10 KB indices)
constexpr int SPMV_U = 8;       // gathers in flight per thread per batch

__device__ __forceinline__ int4 ld_stream_i4(const int4 *p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ double2 ld_stream_d2(const double2 *p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

struct SpmvKernelParams {
    const int *row_ptr; const int *col; const double *val;
    int row_begin, row_end;  // rows handled by this launch ...
    int row2_begin, row2_end;  // ... plus an optional second range (the two boundary slabs of a
    int nchunks1;              //     distributed operator in one launch); chunks >= nchunks1 map to it
    const double *x; long long ldx;
    double *y; long long ldy;
    const double *b; long long ldb;
    const double *d;
    int k;
    double *dot_partials;
    // producer-side halo exchange (dist.cu): the first nsig chunks hold every row a peer needs; each consumer warp
    // bumps *sig when its rows of such a chunk are stored, and the concurrently running exchange kernel starts
    // packing when the count is complete
    int nsig;
    unsigned *sig;
    // fused push: the rows of the signalled chunks also go straight into the neighbours' receive buffers, and the warp
    // that completes the count (sig_target) publishes the epoch in their flag slots -- no pack kernel at all
    const P2PPlanDev *push_plan;
    const int *push_map;
    int push_lo;
    unsigned sig_target;
    unsigned long long *trace;  // in-kernel timeline (CTA 0 stamps its begin / end), nullptr when off
    int trace_id;
};



// End of a signalled chunk (warp-uniform call): count it; with a push plan, first store this row's value into the
// neighbour that needs it, and let the warp that completes the count publish the exchange.  Out of line: the rare
// path must not cost the row walk any registers.
__device__ __noinline__ void chunk_signal(const P2PPlanDev *pl, const int *push_map, int push_lo, int push_hi, unsigned *sig,
                                          unsigned sig_target, const double *y, int row, bool owner) {
    if (pl != nullptr) {
        // a signalled chunk may reach past the boundary rows (chunks hold 256 / tpr rows): only rows outside [push_lo, push_hi)
        // have an entry in the push map
        if (owner && (row < push_lo || row >= push_hi)) {
            const int mi = row < push_lo ? row : row - push_hi + push_lo;
            const int m = __ldg(push_map + mi);
            if (m >= 0) {
                const unsigned long long e = *pl->epoch + 1ull;
                pl->rdst[e & 1ull][m >> PUSH_SLOT_BITS][m & ((1 << PUSH_SLOT_BITS) - 1)] = y[row];
            }
        }
        __syncwarp();
        __threadfence_system();
        if ((threadIdx.x & 31) == 0) {
            const unsigned old = atomicAdd(sig, 1u);
            if (old + 1u == sig_target) {  // every boundary row is stored here and pushed to its neighbour
                __threadfence_system();
                const unsigned long long e = *pl->epoch + 1ull;
                for (int nb = 0; nb < pl->nnb; ++nb) st_release_sys(pl->rflag[nb], e);
                *sig = 0u;
            }
        }
        return;
    }
    __syncwarp();
    __threadfence();
    if ((threadIdx.x & 31) == 0) atomicAdd(sig, 1u);
}

// chunk index -> [r0, r1) over the (up to two) row ranges of a launch
template <int ROWS>
__device__ __forceinline__ void chunk_rows(const SpmvKernelParams &p, int c, int &r0, int &r1) {
    if (c < p.nchunks1) { r0 = p.row_begin + c * ROWS; r1 = min(r0 + ROWS, p.row_end); }
    else { r0 = p.row2_begin + (c - p.nchunks1) * ROWS; r1 = min(r0 + ROWS, p.row2_end); }
}

// Epilogue operands are loaded *before* the row's gathers (and before the stage wait) so their
// latency overlaps the streaming of A instead of being exposed once per chunk.
struct EpiOps { double b, d, x, y; };

template <int EPI, bool DOT>
__device__ __forceinline__ EpiOps epi_prefetch(const SpmvKernelParams &p, const double *__restrict__ xc, const double *yc, int c,
                                               int row, bool on) {
    EpiOps o{0.0, 0.0, 0.0, 0.0};
    if (on) {
        // b, d and y are touched once per sweep: streaming (evict-first) so they do not push the
        // x lines that the gathers reuse out of L1/L2
        if (EPI == EPI_RESID || EPI == EPI_SMOOTH) o.b = __ldcs(p.b + (long long)c * p.ldb + row);
        if (EPI == EPI_SMOOTH || EPI == EPI_SI || EPI == EPI_EPROP) o.d = __ldcs(p.d + row);
        if (EPI == EPI_SMOOTH || EPI == EPI_SI || EPI == EPI_EPROP || DOT) o.x = xc[row];
        if (EPI == EPI_ADD) o.y = __ldcs(yc + row);
    }
    return o;
}

template <int EPI, bool DOT>
__device__ __forceinline__ void epi_store(double *yc, int row, double s, const EpiOps &o, double &dot_acc) {
    if (EPI == EPI_SPMV) {
        yc[row] = s;
    } else if (EPI == EPI_RESID) {
        yc[row] = o.b - s;                       // multigrid.rs:341-342
    } else if (EPI == EPI_SMOOTH) {
        const double r = o.b - s;                // x' = x + d .* (b - A x)   (multigrid.rs:419-422)
        const double xn = o.x + o.d * r;
        yc[row] = xn;
        if (DOT) dot_acc += o.b * xn;            // b . x': PCG's r.z when this is the cycle's last sweep (b = r, x' = z)
    } else if (EPI == EPI_ADD) {
        yc[row] = o.y + s;                       // multigrid.rs:349-350
    } else if (EPI == EPI_EPROP) {
        yc[row] = o.x - o.d * s;                 // ErrorPropogator: out = x - M^-1 (A x)   (adaptivity.rs:191-198)
    } else {
        yc[row] = o.x + o.d * (o.x - s);         // EPI_SI: x' = x + d .* (x - A x)   (smoothers.rs:153-156, literal)
    }
    if (DOT && EPI != EPI_SMOOTH) dot_acc += o.x * s;
}

// ---- column-blocked row walk for k > 1 (K2, SpMM) -------------------------------------------------
// One pass over a row's staged (col, val) pairs feeds CB right-hand sides: CB x SPMV_U gathers are in
// flight per thread instead of SPMV_U, and the walk is repeated k / CB times instead of k times.  With one
// Measured on G7(128^3), E apply at k = 64: CB = 1: 1.27 ms, CB = 2 (72 registers, 3 CTAs/SM): 1.13 ms,
// CB = 4 (96 registers + spills, 2 CTAs/SM): 1.93 ms -- CB = 2 is what ships (profiles/r1_nearnull.md).
// Per column the products are summed in the same order as in the single-column walk: same bits.
template <int TPR, int EPI, int CB, bool STAGED>
__device__ __forceinline__ void spmm_cols(const SpmvKernelParams &p, int col0, int ncb, const double *__restrict__ vals,
                                          const int *__restrict__ cols, int qa, int qe, int row, bool store) {
    constexpr int GPW = 32 / TPR;
    const double *xc[CB];
    EpiOps ops[CB];
    double s[CB];
#pragma unroll
    for (int c = 0; c < CB; ++c) {
        const int col = col0 + (c < ncb ? c : 0);
        xc[c] = p.x + (long long)col * p.ldx;
        ops[c] = epi_prefetch<EPI, false>(p, xc[c], p.y + (long long)col * p.ldy, col, row, store && c < ncb);
        s[c] = 0.0;
    }
    for (int q = qa; q < qe; q += TPR * SPMV_U) {
        double pr[CB][SPMV_U];
#pragma unroll
        for (int u = 0; u < SPMV_U; ++u) {
            const int qq = q + u * TPR;
            double v = 0.0;
            int j = 0;
            const bool on = qq < qe;
            if (on) { v = STAGED ? vals[qq] : __ldg(vals + qq); j = STAGED ? cols[qq] : __ldg(cols + qq); }
#pragma unroll
            for (int c = 0; c < CB; ++c) pr[c][u] = on ? v * __ldg(xc[c] + j) : 0.0;
        }
#pragma unroll
        for (int c = 0; c < CB; ++c)
#pragma unroll
            for (int u = 0; u < SPMV_U; ++u) s[c] += pr[c][u];
    }
    double unused = 0.0;
#pragma unroll
    for (int c = 0; c < CB; ++c) {
        if (TPR > 1) {
#pragma unroll
            for (int o = 16; o >= GPW; o >>= 1) s[c] += __shfl_xor_sync(0xffffffffu, s[c], o);
        }
        if (store && c < ncb) epi_store<EPI, false>(p.y + (long long)(col0 + c) * p.ldy, row, s[c], ops[c], unused);
    }
}

template <int TPR, int EPI, bool DOT, int CB>
__global__ void __launch_bounds__(SPMV_THREADS, CB == 1 ? 4 : 3) spmv_kernel(const SpmvKernelParams p) {
    constexpr int ROWS = SPMV_THREADS / TPR;
    __shared__ __align__(16) double s_val[SPMV_CAP];
    __shared__ __align__(16) int s_col[SPMV_CAP];
    __shared__ double s_red[SPMV_THREADS / 32];

    const int tid = threadIdx.x;
    if (p.trace && tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) trace_rec(p.trace, TR_BEGIN, p.trace_id);
    int r0, r1;
    chunk_rows<ROWS>(p, blockIdx.x, r0, r1);
    const int q0 = __ldg(p.row_ptr + r0);
    const int q1 = __ldg(p.row_ptr + r1);
    const int q0a = q0 & ~3;  // 16-byte aligned start for int4 / 32-byte for double2 pairs
    const int cnt = q1 - q0a;
    const bool staged = cnt <= SPMV_CAP;

    // lane mapping inside a warp: the row varies fastest (GPW = 32/TPR adjacent rows), the entry
    // slot slowest, so one gather instruction reads GPW consecutive x entries per slot for a stencil
    // matrix, and the epilogue lanes (slot 0) hold GPW consecutive rows (coalesced vector traffic)
    constexpr int GPW = 32 / TPR;
    const int g = (tid >> 5) * GPW + ((tid & 31) % GPW);
    const int lane = (tid & 31) / GPW;
    const int row = r0 + g;
    const bool active = row < r1;
    int a = 0, e = 0;
    if (active) { a = __ldg(p.row_ptr + row); e = __ldg(p.row_ptr + row + 1); }
    EpiOps ops = epi_prefetch<EPI, DOT>(p, p.x, p.y, 0, row, active && lane == 0);

    if (staged) {
        const int n4 = (cnt + 3) >> 2;
        const int4 *gc = reinterpret_cast<const int4 *>(p.col + q0a);
        const double2 *gv = reinterpret_cast<const double2 *>(p.val + q0a);
        int4 *sc = reinterpret_cast<int4 *>(s_col);
        double2 *sv = reinterpret_cast<double2 *>(s_val);
        // issue all loads of a thread before the stores (memory-level parallelism)
        for (int i = tid; i < n4; i += SPMV_THREADS * 2) {
            const int i2 = i + SPMV_THREADS;
            int4 c0 = ld_stream_i4(gc + i);
            double2 v0 = ld_stream_d2(gv + 2 * i), v1 = ld_stream_d2(gv + 2 * i + 1);
            int4 c1 = make_int4(0, 0, 0, 0);
            double2 v2 = make_double2(0, 0), v3 = make_double2(0, 0);
            if (i2 < n4) {
                c1 = ld_stream_i4(gc + i2);
                v2 = ld_stream_d2(gv + 2 * i2);
                v3 = ld_stream_d2(gv + 2 * i2 + 1);
            }
            sc[i] = c0; sv[2 * i] = v0; sv[2 * i + 1] = v1;
            if (i2 < n4) { sc[i2] = c1; sv[2 * i2] = v2; sv[2 * i2 + 1] = v3; }
        }
    }
    __syncthreads();

    double dot_acc = 0.0;
    if constexpr (CB > 1) {
        for (int c0 = 0; c0 < p.k; c0 += CB) {
            const int ncb = min(CB, p.k - c0);
            if (staged) spmm_cols<TPR, EPI, CB, true>(p, c0, ncb, s_val, s_col, a - q0a + lane, e - q0a, row, active && lane == 0);
            else spmm_cols<TPR, EPI, CB, false>(p, c0, ncb, p.val, p.col, a + lane, e, row, active && lane == 0);
        }
    } else {
    for (int c = 0; c < p.k; ++c) {
        const double *__restrict__ xc = p.x + (long long)c * p.ldx;
        double *yc = p.y + (long long)c * p.ldy;
        if (c > 0) ops = epi_prefetch<EPI, DOT>(p, xc, yc, c, row, active && lane == 0);
        double s = 0.0;
        if (staged) {
            for (int q = a - q0a + lane; q < e - q0a; q += TPR * SPMV_U) {
                double pr[SPMV_U];
#pragma unroll
                for (int u = 0; u < SPMV_U; ++u) {
                    const int qq = q + u * TPR;
                    pr[u] = 0.0;
                    if (qq < e - q0a) pr[u] = s_val[qq] * __ldg(xc + s_col[qq]);
                }
#pragma unroll
                for (int u = 0; u < SPMV_U; ++u) s += pr[u];
            }
        } else {
            for (int q = a + lane; q < e; q += TPR * SPMV_U) {
                double pr[SPMV_U];
#pragma unroll
                for (int u = 0; u < SPMV_U; ++u) {
                    const int qq = q + u * TPR;
                    pr[u] = 0.0;
                    if (qq < e) pr[u] = __ldg(p.val + qq) * __ldg(xc + __ldg(p.col + qq));
                }
#pragma unroll
                for (int u = 0; u < SPMV_U; ++u) s += pr[u];
            }
        }
        if (TPR > 1) {
#pragma unroll
            for (int o = 16; o >= GPW; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        }
        if (active && lane == 0) epi_store<EPI, DOT>(yc, row, s, ops, dot_acc);
    }
    if ((int)blockIdx.x < p.nsig)  // warp-uniform; signalled launches: row_begin == push_hi
        chunk_signal(p.push_plan, p.push_map, p.push_lo, p.row_begin, p.sig, p.sig_target, p.y, row, active && lane == 0);
    if (DOT) {
        // deterministic CTA reduction of x_i * (A x)_i -> one partial per CTA
        double v = dot_acc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) s_red[tid >> 5] = v;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < SPMV_THREADS / 32; ++w) t += s_red[w];
            p.dot_partials[blockIdx.x] = t;
        }
    }
    }  // CB == 1
    if (p.trace && tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) trace_rec(p.trace, TR_END, p.trace_id);
}

// ===================================================================================================
// Variant 2: persistent, warp-specialised, TMA-fed pipeline.
//   * grid = (resident CTAs per SM) x (SM count); every CTA walks chunks c = blockIdx.x, +gridDim.x, ...
//   * measured configuration sweep on B200 (profiles/r1_kernel_ab.md): 4 CTAs/SM x 2 stages x 2304
//     non-zeros beats 3 x 3 x 2048 by 10-15 % on the fused variants -- consumer warps per SM matter
//     more than pipeline depth; 5 CTAs/SM (40 registers, spills) and 2 CTAs/SM are both clearly worse.
//   * warp 8 is the producer: one elected lane turns each chunk into two bulk async copies
//     (cp.async.bulk global -> shared, SASS UBLKCP) of the chunk's col[] and val[] slices into a
//     TMA_STAGES-deep ring, completion signalled on an mbarrier with expect_tx;
//   * warps 0..7 are consumers: wait for the stage's "full" barrier, walk their rows out of shared
//     memory exactly as variant 1 (same summation order => same bits), then release the stage on
//     its "empty" barrier.  No __syncthreads in the steady state, no register staging of A.
// ===================================================================================================
#ifndef FAMG_TMA_STAGES
#define FAMG_TMA_STAGES 2
#endif
#ifndef FAMG_TMA_CAP
#define FAMG_TMA_CAP 2304
#endif
#ifndef FAMG_TMA_CTAS
#define FAMG_TMA_CTAS 4
#endif
constexpr int TMA_STAGES = FAMG_TMA_STAGES;
constexpr int TMA_CAP = FAMG_TMA_CAP;  // non-zeros per stage (12 bytes each: f64 value + i32 index)
constexpr int TMA_CTAS = FAMG_TMA_CTAS;  // resident CTAs per SM the grid is sized for
constexpr int TMA_CONSUMERS = 256;
constexpr int TMA_THREADS = TMA_CONSUMERS + 32;
constexpr int TMA_SMEM = TMA_STAGES * TMA_CAP * 12 + 2 * TMA_STAGES * 8 + 8 * 8;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int TPR, int EPI, bool DOT, int CB>
__global__ void __launch_bounds__(TMA_THREADS, CB == 1 ? TMA_CTAS : 3) spmv_tma_kernel(const SpmvKernelParams p, const int nchunks) {
    constexpr int ROWS = TMA_CONSUMERS / TPR;
    extern __shared__ __align__(128) unsigned char smem[];
    double *s_val = reinterpret_cast<double *>(smem);
    int *s_col = reinterpret_cast<int *>(smem + (size_t)TMA_STAGES * TMA_CAP * 8);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)TMA_STAGES * TMA_CAP * 12);
    uint64_t *empty = full + TMA_STAGES;
    double *s_red = reinterpret_cast<double *>(empty + TMA_STAGES);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    if (p.trace && tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) trace_rec(p.trace, TR_BEGIN, p.trace_id);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < TMA_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], TMA_CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == TMA_CONSUMERS / 32) {
        // ------------------------------------------------------------------ producer
        if ((tid & 31) == 0) {
            int stage = 0, uses = 0;
            uint32_t ephase = 0;
            for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
                int r0, r1;
                chunk_rows<ROWS>(p, c, r0, r1);
                const int q0a = __ldg(p.row_ptr + r0) & ~3;
                const int cnt = __ldg(p.row_ptr + r1) - q0a;
                if (cnt <= 0 || cnt > TMA_CAP) continue;  // consumers take the direct path
                if (uses >= TMA_STAGES) { mbar_wait(&empty[stage], (ephase >> stage) & 1u); ephase ^= 1u << stage; }
                const uint32_t n4 = (uint32_t)(cnt + 3) >> 2;
                mbar_expect_tx(&full[stage], n4 * 48u);
                bulk_g2s(s_col + stage * TMA_CAP, p.col + q0a, n4 * 16u, &full[stage]);
                bulk_g2s(s_val + stage * TMA_CAP, p.val + q0a, n4 * 32u, &full[stage]);
                ++uses;
                stage = stage + 1 == TMA_STAGES ? 0 : stage + 1;
            }
        }
        return;
    }
    // ---------------------------------------------------------------------- consumers
    constexpr int GPW = 32 / TPR;  // row groups per warp; rows fastest, entry slot slowest (see variant 1)
    const int g = (tid >> 5) * GPW + ((tid & 31) % GPW);
    const int lane = (tid & 31) / GPW;
    int stage = 0;
    uint32_t fphase = 0;
    double dot_acc = 0.0;
    int c = blockIdx.x;
    int q0 = 0, q1 = 0, a = 0, e = 0;
    if (c < nchunks) {
        int r0, r1;
        chunk_rows<ROWS>(p, c, r0, r1);
        q0 = __ldg(p.row_ptr + r0); q1 = __ldg(p.row_ptr + r1);
        if (r0 + g < r1) { a = __ldg(p.row_ptr + r0 + g); e = __ldg(p.row_ptr + r0 + g + 1); }
    }
    while (c < nchunks) {
        int r0, r1;
        chunk_rows<ROWS>(p, c, r0, r1);
        const int row = r0 + g;
        const bool active = row < r1;
        // prefetch the next chunk's row pointers (consumed one iteration later)
        const int cn = c + gridDim.x;
        int nq0 = 0, nq1 = 0, na = 0, ne = 0;
        if (cn < nchunks) {
            int s0, s1;
            chunk_rows<ROWS>(p, cn, s0, s1);
            nq0 = __ldg(p.row_ptr + s0); nq1 = __ldg(p.row_ptr + s1);
            if (s0 + g < s1) { na = __ldg(p.row_ptr + s0 + g); ne = __ldg(p.row_ptr + s0 + g + 1); }
        }
        const int q0a = q0 & ~3;
        const int cnt = q1 - q0a;
        const bool staged = cnt > 0 && cnt <= TMA_CAP;
        const double *sv = s_val + stage * TMA_CAP;
        const int *sc = s_col + stage * TMA_CAP;
        if constexpr (CB > 1) {
            if (staged) { mbar_wait(&full[stage], (fphase >> stage) & 1u); fphase ^= 1u << stage; }
            for (int c0 = 0; c0 < p.k; c0 += CB) {
                const int ncb = min(CB, p.k - c0);
                if (staged) spmm_cols<TPR, EPI, CB, true>(p, c0, ncb, sv, sc, a - q0a + lane, e - q0a, row, active && lane == 0);
                else spmm_cols<TPR, EPI, CB, false>(p, c0, ncb, p.val, p.col, a + lane, e, row, active && lane == 0);
            }
        } else
        for (int col = 0; col < p.k; ++col) {
            const double *__restrict__ xc = p.x + (long long)col * p.ldx;
            double *yc = p.y + (long long)col * p.ldy;
            const EpiOps ops = epi_prefetch<EPI, DOT>(p, xc, yc, col, row, active && lane == 0);
            if (col == 0 && staged) { mbar_wait(&full[stage], (fphase >> stage) & 1u); fphase ^= 1u << stage; }
            double s = 0.0;
            if (staged) {
                for (int q = a - q0a + lane; q < e - q0a; q += TPR * SPMV_U) {
                    double pr[SPMV_U];
#pragma unroll
                    for (int u = 0; u < SPMV_U; ++u) {
                        const int qq = q + u * TPR;
                        pr[u] = 0.0;
                        if (qq < e - q0a) pr[u] = sv[qq] * __ldg(xc + sc[qq]);
                    }
#pragma unroll
                    for (int u = 0; u < SPMV_U; ++u) s += pr[u];
                }
            } else {
                for (int q = a + lane; q < e; q += TPR * SPMV_U) {
                    double pr[SPMV_U];
#pragma unroll
                    for (int u = 0; u < SPMV_U; ++u) {
                        const int qq = q + u * TPR;
                        pr[u] = 0.0;
                        if (qq < e) pr[u] = __ldg(p.val + qq) * __ldg(xc + __ldg(p.col + qq));
                    }
#pragma unroll
                    for (int u = 0; u < SPMV_U; ++u) s += pr[u];
                }
            }
            if (TPR > 1) {
#pragma unroll
                for (int o = 16; o >= GPW; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            }
            if (active && lane == 0) epi_store<EPI, DOT>(yc, row, s, ops, dot_acc);
        }
        if (c < p.nsig)  // warp-uniform: this chunk's rows are awaited by the peers (signalled launches: row_begin == push_hi)
            chunk_signal(p.push_plan, p.push_map, p.push_lo, p.row_begin, p.sig, p.sig_target, p.y, row, active && lane == 0);
        if (staged) {
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty[stage]);
            stage = stage + 1 == TMA_STAGES ? 0 : stage + 1;
        }
        c = cn; q0 = nq0; q1 = nq1; a = na; e = ne;
    }
    if (p.trace && tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) trace_rec(p.trace, TR_END, p.trace_id);
    if (DOT) {
        double v = dot_acc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) s_red[warp] = v;
        asm volatile("bar.sync 1, %0;" ::"n"(TMA_CONSUMERS) : "memory");
        if (tid == 0) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < TMA_CONSUMERS / 32; ++w) t += s_red[w];
            p.dot_partials[blockIdx.x] = t;
        }
    }
}

template <int TPR, int EPI, bool DOT, int CB>
static famg_status launch_one(SpmvKernelParams kp, int variant, int nrows1, int nrows2, int num_sms, int reserve_ctas, cudaStream_t st,
                              int *grid_out) {
    constexpr int ROWS = 256 / TPR;  // both variants use 256 row-walking threads
    kp.nchunks1 = (int)ceil_div(nrows1, ROWS);
    // signalled launches run [push_hi, n) as the first range and [0, push_hi) as the second; kp.nsig arrives holding
    // push_lo: the signalled chunks are the whole first range and the chunks covering [0, push_lo) of the second
    if (kp.sig) kp.nsig = kp.nchunks1 + (int)ceil_div(kp.nsig, ROWS); else kp.nsig = 0;
    const int nchunks = kp.nchunks1 + (int)ceil_div(nrows2, ROWS);
    if (variant == 2) {
        constexpr int CTAS = CB == 1 ? TMA_CTAS : 3;  // resident CTAs per SM (register budget of the column block)
        const int grid = std::max(1, std::min(nchunks, CTAS * num_sms - reserve_ctas));
        // the > 48 KB opt-in is a per-device attribute: one bit per device and template instance
        static std::atomic<uint64_t> configured{0};
        int dev = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        if (!(configured.load(std::memory_order_relaxed) >> (dev & 63) & 1ull)) {
            CUDA_TRY(cudaFuncSetAttribute(spmv_tma_kernel<TPR, EPI, DOT, CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM));
            configured.fetch_or(1ull << (dev & 63), std::memory_order_relaxed);
        }
        spmv_tma_kernel<TPR, EPI, DOT, CB><<<grid, TMA_THREADS, TMA_SMEM, st>>>(kp, nchunks);
        *grid_out = grid;
    } else {
        spmv_kernel<TPR, EPI, DOT, CB><<<nchunks, SPMV_THREADS, 0, st>>>(kp);
        *grid_out = nchunks;
    }
    return FAMG_OK;
}

template <int TPR, int EPI>
static famg_status launch_cb(const SpmvKernelParams &kp, int cb, int variant, int nrows, int nrows2, int sms, int reserve, cudaStream_t st, int *grid) {
    if (cb >= 2) return launch_one<TPR, EPI, false, 2>(kp, variant, nrows, nrows2, sms, reserve, st, grid);
    return launch_one<TPR, EPI, false, 1>(kp, variant, nrows, nrows2, sms, reserve, st, grid);
}

template <int TPR>
static famg_status launch_tpr(const SpmvKernelParams &kp, int epi, bool dot, int cb, int variant, int nrows, int nrows2, int sms, int reserve, cudaStream_t st, int *grid) {
    switch (epi) {
        case EPI_SPMV:
            return dot ? launch_one<TPR, EPI_SPMV, true, 1>(kp, variant, nrows, nrows2, sms, reserve, st, grid)
                       : launch_cb<TPR, EPI_SPMV>(kp, cb, variant, nrows, nrows2, sms, reserve, st, grid);
        case EPI_RESID: return launch_cb<TPR, EPI_RESID>(kp, cb, variant, nrows, nrows2, sms, reserve, st, grid);
        case EPI_SMOOTH:
            return dot ? launch_one<TPR, EPI_SMOOTH, true, 1>(kp, variant, nrows, nrows2, sms, reserve, st, grid)
                       : launch_cb<TPR, EPI_SMOOTH>(kp, cb, variant, nrows, nrows2, sms, reserve, st, grid);
        case EPI_ADD: return launch_cb<TPR, EPI_ADD>(kp, cb, variant, nrows, nrows2, sms, reserve, st, grid);
        case EPI_EPROP: return launch_cb<TPR, EPI_EPROP>(kp, cb, variant, nrows, nrows2, sms, reserve, st, grid);
        default: return launch_cb<TPR, EPI_SI>(kp, cb, variant, nrows, nrows2, sms, reserve, st, grid);
    }
}

famg_status spmv_launch(const SpmvArgs &args, int *num_ctas) {
    const famg_csr *a = args.a;
    famg_ctx *ctx = a->ctx;
    cudaStream_t st = args.stream ? args.stream : ctx->stream;
    int row_begin = args.row_begin;
    int row_end = args.row_end < 0 ? (int)a->nrows : args.row_end;
    if (num_ctas) *num_ctas = 0;
    if ((row_end <= row_begin && args.row2_end <= args.row2_begin) || args.k <= 0) return FAMG_OK;
    if (row_end < row_begin) row_end = row_begin;
    if ((args.epi == EPI_RESID || args.epi == EPI_SMOOTH) && !args.b) FAMG_FAIL(FAMG_ERR_INVALID, "spmv: missing rhs");
    if ((args.epi == EPI_SMOOTH || args.epi == EPI_SI || args.epi == EPI_EPROP) && (!args.d || args.y == args.x))
        FAMG_FAIL(FAMG_ERR_INVALID, "spmv: smoother sweep needs a diagonal and distinct in/out vectors");
    if (args.dot_partials && (args.k != 1 || (args.epi != EPI_SPMV && args.epi != EPI_SMOOTH)))
        FAMG_FAIL(FAMG_ERR_INVALID, "spmv: fused dot products exist for y = A x (x . A x) and the smoother sweep (b . x'), k == 1");
    SpmvKernelParams kp;
    kp.row_ptr = a->row_ptr; kp.col = a->col; kp.val = a->val;
    kp.row_begin = row_begin; kp.row_end = row_end;
    kp.row2_begin = args.row2_begin; kp.row2_end = std::max(args.row2_end, args.row2_begin); kp.nchunks1 = 0;
    kp.x = args.x; kp.ldx = args.ldx; kp.y = args.y; kp.ldy = args.ldy;
    kp.b = args.b; kp.ldb = args.ldb; kp.d = args.d; kp.k = args.k;
    kp.dot_partials = args.dot_partials;
    kp.sig = args.sig; kp.nsig = args.sig ? args.sig_lo : 0;
    kp.push_plan = args.sig ? (const P2PPlanDev *)args.push_plan : nullptr; kp.push_map = args.push_map; kp.push_lo = args.sig_lo;
    kp.sig_target = args.sig ? spmv_signal_target(a, args.sig_lo, args.sig_hi) : 0u;
    kp.trace = ctx->d_trace;
    kp.trace_id = ctx->d_trace ? trace_new_id(ctx, "spmv epi=%d rows=%lld nnz=%lld tpr=%d%s", args.epi, (long long)a->nrows, (long long)a->nnz,
                                              a->tpr, args.sig ? " signalled" : "") : -1;
    if (args.sig) {  // boundary-first row order (see SpmvArgs::sig)
        if (args.k != 1 || args.sig_lo < 0 || args.sig_hi < args.sig_lo || args.sig_hi > (int)a->nrows || args.row_begin != 0 || args.row_end >= 0 ||
            args.row2_end > args.row2_begin)
            FAMG_FAIL(FAMG_ERR_INVALID, "internal: signalled launches cover all rows of a single right-hand side");
        kp.row_begin = row_begin = args.sig_hi; kp.row_end = row_end = (int)a->nrows;
        kp.row2_begin = 0; kp.row2_end = args.sig_hi;
    }
    const int tpr = a->tpr;
    const int nrows = row_end - row_begin;
    const int nrows2 = kp.row2_end - kp.row2_begin;
    const bool dot = args.dot_partials != nullptr;
    // small operators (coarse levels) do not fill a persistent grid: keep the one-chunk-per-CTA kernel
    const int variant = (ctx->spmv_variant == 2 && nrows + nrows2 >= ctx->tma_min_rows) ? 2 : 1;
    // right-hand sides walked per pass over a row (K2): 1 for k == 1 (the cycle / PCG path is untouched)
    const int cb = args.k >= 2 ? std::min(ctx->spmm_cb, 2) : 1;
    int grid = 0;
    famg_status stt;
    switch (tpr) {
        case 1: stt = launch_tpr<1>(kp, args.epi, dot, cb, variant, nrows, nrows2, ctx->num_sms, args.reserve_ctas, st, &grid); break;
        case 2: stt = launch_tpr<2>(kp, args.epi, dot, cb, variant, nrows, nrows2, ctx->num_sms, args.reserve_ctas, st, &grid); break;
        case 4: stt = launch_tpr<4>(kp, args.epi, dot, cb, variant, nrows, nrows2, ctx->num_sms, args.reserve_ctas, st, &grid); break;
        case 8: stt = launch_tpr<8>(kp, args.epi, dot, cb, variant, nrows, nrows2, ctx->num_sms, args.reserve_ctas, st, &grid); break;
        case 16: stt = launch_tpr<16>(kp, args.epi, dot, cb, variant, nrows, nrows2, ctx->num_sms, args.reserve_ctas, st, &grid); break;
        default: stt = launch_tpr<32>(kp, args.epi, dot, cb, variant, nrows, nrows2, ctx->num_sms, args.reserve_ctas, st, &grid); break;
    }
    FAMG_TRY(stt);
    count_launch(ctx);
    KERNEL_CHECK();
    if (num_ctas) *num_ctas = grid;
    return FAMG_OK;
}

}  // namespace famg

using namespace famg;

static famg_status check_apply(const famg_csr *a, const famg_vec *out, const famg_vec *rhs) {
    if (!a || !out || !rhs) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    if (rhs->nrows != a->ncols || out->nrows != a->nrows || out->ncols != rhs->ncols)
        FAMG_FAIL(FAMG_ERR_INVALID, "spmm shape mismatch: A %lldx%lld, rhs %lldx%lld, out %lldx%lld", (long long)a->nrows,
                  (long long)a->ncols, (long long)rhs->nrows, (long long)rhs->ncols, (long long)out->nrows, (long long)out->ncols);
    if (out->p == rhs->p) FAMG_FAIL(FAMG_ERR_INVALID, "out aliases rhs");
    return FAMG_OK;
}

extern "C" {

famg_status famg_spmm_dev(const famg_csr *a, famg_vec *out, const famg_vec *rhs) {
    FAMG_TRY(check_apply(a, out, rhs));
    CUDA_TRY(cudaSetDevice(a->ctx->device));
    SpmvArgs s; s.a = a; s.epi = EPI_SPMV; s.x = rhs->p; s.ldx = rhs->ld; s.y = out->p; s.ldy = out->ld; s.k = (int)rhs->ncols;
    return spmv_launch(s);
}

famg_status famg_residual_dev(const famg_csr *a, famg_vec *out, const famg_vec *b, const famg_vec *x) {
    FAMG_TRY(check_apply(a, out, x));
    if (!b || b->nrows != a->nrows || b->ncols != x->ncols) FAMG_FAIL(FAMG_ERR_INVALID, "residual: rhs shape mismatch");
    CUDA_TRY(cudaSetDevice(a->ctx->device));
    SpmvArgs s; s.a = a; s.epi = EPI_RESID; s.x = x->p; s.ldx = x->ld; s.y = out->p; s.ldy = out->ld;
    s.b = b->p; s.ldb = b->ld; s.k = (int)x->ncols;
    return spmv_launch(s);
}

famg_status famg_spmm_add_dev(const famg_csr *a, famg_vec *y, const famg_vec *x) {
    FAMG_TRY(check_apply(a, y, x));
    CUDA_TRY(cudaSetDevice(a->ctx->device));
    SpmvArgs s; s.a = a; s.epi = EPI_ADD; s.x = x->p; s.ldx = x->ld; s.y = y->p; s.ldy = y->ld; s.k = (int)x->ncols;
    return spmv_launch(s);
}

// LinOp::apply(out: MatMut, rhs: MatRef) with host buffers: stage in, apply, stage out.
famg_status famg_spmm(const famg_csr *a, double *out, int64_t ld_out, const double *rhs, int64_t ld_rhs, int64_t k) {
    if (!a || !out || !rhs || k < 0) FAMG_FAIL(FAMG_ERR_INVALID, "bad argument");
    if (ld_out < a->nrows || ld_rhs < a->ncols) FAMG_FAIL(FAMG_ERR_INVALID, "leading dimension too small");
    if (k == 0) return FAMG_OK;
    famg_vec *x = nullptr, *y = nullptr;
    FAMG_TRY(famg_vec_create(a->ctx, a->ncols, k, &x));
    famg_status st = famg_vec_create(a->ctx, a->nrows, k, &y);
    if (st == FAMG_OK) st = famg_vec_upload(x, rhs, ld_rhs);
    if (st == FAMG_OK) st = famg_spmm_dev(a, y, x);
    if (st == FAMG_OK) st = famg_vec_download(y, out, ld_out);
    famg_vec_destroy(x); famg_vec_destroy(y);
    return st;
}

famg_status famg_time_kernel(const famg_csr *a, int which, int reps, int warmup, float *ms_avg) {
    if (!a || !ms_avg || reps <= 0 || which < 0 || which > 3) FAMG_FAIL(FAMG_ERR_INVALID, "bad argument");
    if (which == 2 && a->nrows != a->ncols) FAMG_FAIL(FAMG_ERR_INVALID, "the smoother sweep needs a square operator");
    famg_ctx *ctx = a->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    famg_vec *x = nullptr, *y = nullptr, *b = nullptr, *d = nullptr;
    FAMG_TRY(famg_vec_create(ctx, a->ncols, 1, &x));
    famg_status st = famg_vec_create(ctx, a->nrows, 1, &y);
    if (st == FAMG_OK) st = famg_vec_create(ctx, a->nrows, 1, &b);
    if (st == FAMG_OK) st = famg_vec_create(ctx, a->nrows, 1, &d);
    if (st == FAMG_OK) st = famg_vec_fill(x, 1.0);
    if (st == FAMG_OK) st = famg_vec_fill(b, 1.0);
    if (st == FAMG_OK) st = famg_vec_fill(d, 0.125);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (st == FAMG_OK) {
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        SpmvArgs s; s.a = a; s.x = x->p; s.ldx = x->ld; s.y = y->p; s.ldy = y->ld; s.b = b->p; s.ldb = b->ld; s.d = d->p; s.k = 1;
        s.epi = which == 0 ? EPI_SPMV : which == 1 ? EPI_RESID : which == 2 ? EPI_SMOOTH : EPI_ADD;
        for (int i = 0; i < warmup && st == FAMG_OK; ++i) st = spmv_launch(s);
        cudaEventRecord(e0, ctx->stream);
        for (int i = 0; i < reps && st == FAMG_OK; ++i) st = spmv_launch(s);
        cudaEventRecord(e1, ctx->stream);
        cudaError_t e = cudaEventSynchronize(e1);
        float ms = 0;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (e != cudaSuccess && st == FAMG_OK) { set_error("timing failed: %s", cudaGetErrorString(e)); st = FAMG_ERR_CUDA; }
        *ms_avg = ms / (float)reps;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    famg_vec_destroy(x); famg_vec_destroy(y); famg_vec_destroy(b); famg_vec_destroy(d);
    return st;
}

}  // extern "C"
