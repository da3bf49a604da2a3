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
//   pass 0  ub_i = sum_k nnz(B_k)                 upper bound of the row length
//   pass 1  rows binned by ub; insert the candidate columns into a key-only table sized from ub
//           -> exact nnz(C_i); scan -> row_ptr
//   pass 2  rows re-binned by their *exact* length (tables sized 2x the distinct count, not the
//           often 10x larger ub); re-insert, compact, bitonic-sort the distinct columns (sorted
//           pattern, bit-exact with a CPU Gustavson), then accumulate: k runs *sequentially*
//           (ascending), the lanes of the row's thread group take the entries of B_k in parallel
//           and locate their slot by binary search.  Within one k every j is unique, so there are
//           no write conflicts and no atomics on values: the result is deterministic and has the
//           reference's summation order.
//   classes one thread per row (<= 32 products), one warp per row (tables of 64..4096 slots in shared memory), one CTA
//           per row (up to 192 KB of shared memory; one f64 accumulator per column when B is narrow), per-CTA tables in
//           global memory for anything larger -- and the window class (sg_window_kernel): rows with thousands of products
//           whose candidate columns fit a window of <= 11 264 columns are computed ONCE, by one warp per row, in the count pass.
// The prolongator-smoothing epilogue  S_i <- -(w/a_ii) S_i + P_i  is fused into pass 2.
#include <cuda_pipeline.h>

#include "common.cuh"

namespace famg {

constexpr double JACOBI_WEIGHT_BLOCK = 0.66;  // hard-coded in block_jacobi (interpolation/mod.rs:1009)

struct SgMat { const int *rp; const int *col; const double *val; };

struct SgEpilogue {
    int enabled;          // 0 none | 1 S_i <- -(w/a_ii) S_i + P_i (smooth_interpolation) | 2 S_i <- S_i + P_i (block_jacobi)
                          // | 3 S_i <- -S_i + P_i (smooth_p)
    double omega;
    SgMat p;              // P (same rows as C)
    int *error_flag;      // 1: missing / tiny diagonal, 2: pattern(P) not in pattern(A P)
};

// epilogue scaling of one accumulated value (see SgEpilogue::enabled); negation is exact, so kind 3
// equals the reference's `m_inv * (-(A P))` bit for bit
__device__ __forceinline__ double sg_epi_scale(int kind, double v, double scalar) {
    return kind == 1 ? v * -scalar : kind == 3 ? -v : v;
}

__device__ __forceinline__ unsigned sg_hash(int j, int hmask) { return ((unsigned)j * 2654435761u >> 8) & (unsigned)hmask; }

template <int GROUP>
__device__ __forceinline__ void group_sync() {
    if (GROUP == 32) __syncwarp(); else __syncthreads();
}

// Insert every candidate column of row i of A*B into `table` (size hmask+1, cleared here).
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

// ---- staged accumulation for rows that own a whole CTA ------------------------------------------
// The plain k loop of sg_row_fill costs one chain of dependent global loads (a.col -> b.rp -> b.col)
// plus a CTA barrier per k: ~1 us per k whatever the length of B_k, which is what made the
// coarse-level products (hundreds to thousands of k per row) the slowest part of a hierarchy build.
// Here the entries of a batch of k's (up to SG_SK segments / SG_SE entries) are brought in by all
// 256 threads at once (the loads of a batch are independent), their output slots are found by a flat,
// perfectly balanced binary-search pass, and only the short in-order accumulation
// vals[slot] += a_ik * b_kj  runs k after k out of shared memory -- by one warp with __syncwarp
// between k's when the segments are short, by the whole CTA when they are long.  The order of
// the adds per output entry is still ascending k: same bits as before.
constexpr int SG_SK = 64;    // segments (k's) per batch
constexpr int SG_SE = 2048;  // entries per batch

struct SgStage {
    double bv[SG_SE];
    double av[SG_SK];
    int slot[SG_SE];
    int b0[SG_SK], len[SG_SK], off[SG_SK + 1];
    int nk, partial;
};

// column -> output slot.  slot16 mirrors the key table of the insertion pass: position h holds the rank of the key that
// landed there among the sorted distinct columns, so a lookup re-walks the key's own probe sequence (1-2 steps at load
// <= 0.5) and recognises it by list[slot] == j -- two shared-memory reads per step instead of the 7-9 dependent reads of a
// bisection (45 % of the instructions of the fill kernels in round 1's source-level profile).  Without a slot table
// (global-memory class) the bisection stays.
__device__ __forceinline__ int sg_find_slot(int j, const int *__restrict__ list, int n, const unsigned short *__restrict__ slot16, int hmask) {
    if (slot16 != nullptr) {
        unsigned h = sg_hash(j, hmask);
        while (true) {
            const int sl = slot16[h];
            if (list[sl] == j) return sl;
            h = (h + 1) & (unsigned)hmask;
        }
    }
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (list[mid] < j) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// DENSE: `vals` is indexed by the column itself (ncols(B) accumulators), `touched` records the
// structural pattern; no search pass.
template <bool DENSE>
__device__ void sg_accumulate_staged(int tid, const SgMat &a, const SgMat &b, int a0, int a1, const int *list, int n, double *vals,
                                     SgStage &st, unsigned char *touched = nullptr, const unsigned short *slot16 = nullptr, int hmask = 0) {
    const int warp = tid >> 5, lane = tid & 31;
    int q = a0, poff = 0;  // next k of the row; entries of that k already consumed (a B_k longer than one batch)
    while (q < a1) {
        if (tid < SG_SK) {
            int len = 0, b0 = 0;
            double av = 0.0;
            if (q + tid < a1) {
                const int k = a.col[q + tid];
                av = a.val[q + tid];
                b0 = b.rp[k];
                len = b.rp[k + 1] - b0;
                if (tid == 0) { b0 += poff; len -= poff; }
            }
            st.av[tid] = av; st.b0[tid] = b0; st.len[tid] = len;
        }
        __syncthreads();
        if (tid == 0) {
            const int avail = min(SG_SK, a1 - q);
            int acc = 0, nk = 0, partial = 0;
            while (nk < avail && acc + st.len[nk] <= SG_SE) { st.off[nk] = acc; acc += st.len[nk]; ++nk; }
            if (nk == 0) { st.off[0] = 0; st.len[0] = SG_SE; acc = SG_SE; nk = 1; partial = 1; }  // B_k longer than a batch
            st.off[nk] = acc; st.nk = nk; st.partial = partial;
        }
        __syncthreads();
        const int nk = st.nk, total = st.off[nk];
        for (int kk = warp; kk < nk; kk += 8) {  // one warp per segment: coalesced, independent loads
            const int b0 = st.b0[kk], len = st.len[kk], off = st.off[kk];
            for (int p = lane; p < len; p += 32) {
                const int j = b.col[b0 + p];
                st.slot[off + p] = j; st.bv[off + p] = b.val[b0 + p];
                if (DENSE) touched[j] = 1;  // same value from every writer
            }
        }
        __syncthreads();
        if (!DENSE) {
            for (int t = tid; t < total; t += 256) st.slot[t] = sg_find_slot(st.slot[t], list, n, slot16, hmask);  // column -> output slot
            __syncthreads();
        }
        if (total > 64 * nk) {  // long segments: the whole CTA takes one k at a time
            for (int kk = 0; kk < nk; ++kk) {
                const double av = st.av[kk];
                const int len = st.len[kk], off = st.off[kk];
                for (int p = tid; p < len; p += 256) { const int sl = st.slot[off + p]; vals[sl] = vals[sl] + av * st.bv[off + p]; }
                __syncthreads();
            }
        } else {
            if (warp == 0) {
                for (int kk = 0; kk < nk; ++kk) {
                    const double av = st.av[kk];
                    const int len = st.len[kk], off = st.off[kk];
                    for (int p = lane; p < len; p += 32) { const int sl = st.slot[off + p]; vals[sl] = vals[sl] + av * st.bv[off + p]; }
                    __syncwarp();
                }
            }
            __syncthreads();
        }
        if (st.partial) poff += SG_SE; else { q += nk; poff = 0; }
        __syncthreads();  // st.partial / st.nk are rewritten by the next round
    }
}

// Pass 2 for one row.  table: H = hmask+1 ints (8-byte aligned; reused as the H/2 f64 accumulators
// once the keys have been compacted), list: H/2 ints.  The row has n <= H/2 distinct columns.
template <int GROUP>
__device__ void sg_row_fill(int i, int lane, const SgMat &a, const SgMat &b, int *table, int *list, int hmask, int *cnt,
                            const int *c_rp, int *c_col, double *c_val, const SgEpilogue &ep, SgStage *stage = nullptr,
                            unsigned short *slot16 = nullptr) {
    const int n = sg_row_insert<GROUP>(i, lane, a, b, table, hmask, cnt);
    const int half = (hmask + 1) >> 1;
    // compact the distinct columns, pad with INT_MAX up to the sort width (power of two >= n)
    int width = 1;
    while (width < n) width <<= 1;
    if (lane == 0) *cnt = 0;
    for (int t = lane; t < width; t += GROUP) list[t] = 0x7fffffff;
    group_sync<GROUP>();
    for (int t = lane; t <= hmask; t += GROUP) {
        const int key = table[t];
        if (key != -1) list[atomicAdd(cnt, 1)] = key;
    }
    group_sync<GROUP>();
    // bitonic sort of list[0..width) (keys distinct) -> sorted pattern, identical to a CPU Gustavson
    for (int k = 2; k <= width; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < width; t += GROUP) {
                const int u = t ^ j;
                if (u > t) {
                    const int x = list[t], y = list[u];
                    if ((x > y) == ((t & k) == 0)) { list[t] = y; list[u] = x; }
                }
            }
            group_sync<GROUP>();
        }
    }
    if (slot16 != nullptr) {
        // rank of every key at the position it occupies in the key table (unique per key: no conflicts)
        for (int t = lane; t < n; t += GROUP) {
            const int key = list[t];
            unsigned h = sg_hash(key, hmask);
            while (table[h] != key) h = (h + 1) & (unsigned)hmask;
            slot16[h] = (unsigned short)t;
        }
        group_sync<GROUP>();
    }
    double *vals = reinterpret_cast<double *>(table);  // the keys are in `list` (and their ranks in slot16) now
    (void)half;
    for (int t = lane; t < n; t += GROUP) vals[t] = 0.0;
    group_sync<GROUP>();
    const int a0 = a.rp[i], a1 = a.rp[i + 1];
    if (GROUP == 256 && stage != nullptr) {
        sg_accumulate_staged<false>(lane, a, b, a0, a1, list, n, vals, *stage, nullptr, slot16, hmask);
    } else {
    // numeric: ascending k, lanes over the entries of B_k (each j unique within one k => no races)
    // The loop is latency-bound (a.col[q] -> b.rp[k] -> b.col[p] are dependent loads and every k ends
    // in a group barrier), so the next k's row bounds and this lane's first entry of B_k are fetched
    // one iteration ahead.
    int k_n = 0, b0_n = 0, b1_n = 0, j_n = 0;
    double av_n = 0.0, bv_n = 0.0;
    if (a0 < a1) {
        k_n = a.col[a0]; av_n = a.val[a0]; b0_n = b.rp[k_n]; b1_n = b.rp[k_n + 1];
        if (b0_n + lane < b1_n) { j_n = b.col[b0_n + lane]; bv_n = b.val[b0_n + lane]; }
    }
    for (int q = a0; q < a1; ++q) {
        const double av = av_n;
        const int b0 = b0_n, b1 = b1_n;
        int j = j_n;
        double bv = bv_n;
        if (q + 1 < a1) {  // prefetch iteration q+1
            k_n = a.col[q + 1]; av_n = a.val[q + 1]; b0_n = b.rp[k_n]; b1_n = b.rp[k_n + 1];
            if (b0_n + lane < b1_n) { j_n = b.col[b0_n + lane]; bv_n = b.val[b0_n + lane]; }
        }
        for (int p = b0 + lane; p < b1; p += GROUP) {
            if (p != b0 + lane) { j = b.col[p]; bv = b.val[p]; }
            const int lo = sg_find_slot(j, list, n, slot16, hmask);
            vals[lo] = vals[lo] + av * bv;
        }
        group_sync<GROUP>();
    }
    }
    const int base = c_rp[i];
    if (!ep.enabled) {
        for (int t = lane; t < n; t += GROUP) { c_col[base + t] = list[t]; c_val[base + t] = vals[t]; }
    } else {
        // smooth_interpolation: scalar = w * (1/a_ii); v *= -scalar; then += P_ij where present
        // The diagonal is searched linearly by the whole group: the rows of a distributed slab are
        // renumbered to [owned | ghost] in place and are no longer sorted by (local) column id.
        if (lane == 0) *cnt = -1;
        group_sync<GROUP>();
        for (int q = a0 + lane; q < a1; q += GROUP) if (a.col[q] == i) *cnt = q;
        group_sync<GROUP>();
        const int qd = *cnt;
        const bool found = qd >= 0;
        const double dv = found ? a.val[qd] : 0.0;
        group_sync<GROUP>();
        if (lane == 0 && ep.enabled == 1 && (!found || !(dv > 1e-6))) atomicMax(ep.error_flag, 1);
        const double scalar = ep.omega * (1.0 / dv);
        const int p0 = ep.p.rp[i], p1 = ep.p.rp[i + 1];
        int matched = 0;
        for (int t = lane; t < n; t += GROUP) {
            const int j = list[t];
            double v = sg_epi_scale(ep.enabled, vals[t], scalar);
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
constexpr int SG_NCLS = 8;   // size classes per pass (see sg_plan)
constexpr int SG_TINY = 32;  // rows with at most this many products: one thread per row
constexpr int SG_WARPS = 4;  // rows per CTA in the warp-per-row classes

// ---- window class: rows whose candidate columns all lie in a short range [lo, lo + span) -----------------
// The products of a Galerkin triple product on the coarse levels have hundreds to thousands of B-rows per output row, and
// those rows overlap heavily: with a locality-preserving numbering all of them fall into a window of a few thousand
// columns.  Such rows get one f64 accumulator per column of the window in shared memory (no table, no sort, no search) and
// are computed ONCE: the count pass already does the arithmetic and parks the finished row in a scratch buffer, the fill
// pass only copies it (see sg_window_kernel).
constexpr int WR_SE = 512;          // entries of B staged per chunk (two chunks per warp: one in flight, one being accumulated)
constexpr int WN_SPAN_MAX = 11264;  // widest window: 9 * span + 12 KB of shared memory per row, two rows per SM
constexpr int WN_CLASS = 7, SG_SKIP_CLASS = 5, SG_GLOBAL_CLASS = 6;

struct SgWinPlan {               // inputs / outputs of the classification of window rows (device pointers; bfirst == nullptr: off)
    const int *bfirst, *blast;   // first / last column of every row of B
    const int *unsorted;         // != 0: some row of B is not strictly ascending -> no window rows
    int *lo, *span;              // per row of A
    unsigned long long *bound;   // sum over the window rows of min(ub, span): what the scratch buffer has to hold at most
    int *max_span;
    int min_ub, min_seg;         // rows with more than min_ub products in segments (rows of B) of at least min_seg entries on average
};

// first / last column of every row of B (8 lanes per row) and whether every row is strictly ascending
__global__ void __launch_bounds__(256) sg_bspan_kernel(SgMat b, int nrows, int *__restrict__ bfirst, int *__restrict__ blast,
                                                       int *__restrict__ unsorted) {
    const int r = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 3), l = threadIdx.x & 7;
    if (r >= nrows) return;
    const int b0 = b.rp[r], b1 = b.rp[r + 1];
    bool bad = false;
    for (int p = b0 + l; p + 1 < b1; p += 8) bad |= b.col[p] >= b.col[p + 1];
    if (bad) atomicOr(unsorted, 1);
    if (l == 0) { bfirst[r] = b0 < b1 ? b.col[b0] : 0x7fffffff; blast[r] = b0 < b1 ? b.col[b1 - 1] : -1; }
}

// ub_i = number of products of row i (work); size_i = min(ub_i, ncols(B)) bounds the distinct columns
__global__ void __launch_bounds__(256) sg_ub_kernel(SgMat a, SgMat b, int m, int ncols_b, int *__restrict__ ub, int *__restrict__ size,
                                                    SgWinPlan wp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    long long s = 0;
    int jmin = 0x7fffffff, jmax = -1;
    if (i < m) {
        for (int q = a.rp[i]; q < a.rp[i + 1]; ++q) {
            const int k = a.col[q];
            s += b.rp[k + 1] - b.rp[k];
            if (wp.bfirst) { jmin = min(jmin, wp.bfirst[k]); jmax = max(jmax, wp.blast[k]); }
        }
    }
    const int u = s > 0x3fffffff ? 0x3fffffff : (int)s;
    unsigned wbound = 0, wspan = 0;
    if (i < m) {
        ub[i] = u;
        // rows with a lot of products get a whole CTA in the count pass even when B is narrow
        int sz = u > 3072 ? max(min(u, ncols_b), 3073) : min(u, ncols_b);
        if (wp.bfirst && u > wp.min_ub && (long long)u >= (long long)wp.min_seg * (a.rp[i + 1] - a.rp[i]) && jmax >= jmin &&
            jmax - jmin < WN_SPAN_MAX && *wp.unsorted == 0) {
            const int span = jmax - jmin + 1;
            wp.lo[i] = jmin; wp.span[i] = span;
            wbound = (unsigned)min(u, span); wspan = (unsigned)span;
            sz = -WN_CLASS;
        }
        size[i] = sz;
    }
    if (wp.bfirst) {  // one atomic per warp
        const unsigned tb = __reduce_add_sync(0xffffffffu, wbound), ts = __reduce_max_sync(0xffffffffu, wspan);
        if ((threadIdx.x & 31) == 0 && tb) { atomicAdd(wp.bound, (unsigned long long)tb); atomicMax(wp.max_span, (int)ts); }
    }
}

// pass-2 size: the exact row length, raised for rows with a lot of work per output entry so that
// they get a whole CTA (the k loop is sequential per row; see sg_row_fill)
__global__ void sg_size2_kernel(const int *__restrict__ row_nnz, const int *__restrict__ ub, int m, int *__restrict__ size,
                                const int *__restrict__ saved) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    // window rows: already computed by the count pass (skip), or deferred to the fill pass because the scratch buffer was full
    if (saved && saved[i] != -1) { size[i] = saved[i] >= 0 ? -SG_SKIP_CLASS : -WN_CLASS; return; }
    // ub <= SG_TINY: scalar class (size 0); otherwise at least 1 so that the row lands in a table class
    size[i] = ub[i] <= SG_TINY ? 0 : max(max(row_nnz[i], 1), min(ub[i] >> 4, 2048));
}

struct SgBounds { int limit[SG_NCLS]; };  // class c holds rows with size <= limit[c] (ascending); a negative size names its class

// Rows -> size classes -> `perm` (rows of each class contiguous, ascending inside a class: the
// permutation is deterministic).  Every CTA owns one contiguous chunk of rows; class counts are
// accumulated in shared memory (one global atomic per CTA and class -- a per-warp global atomic on 8
// addresses serialised 500 k atomics for a 16 M-row operator and cost 11 ms per kernel), the per-CTA
// counts are kept so that the binning pass can compute exact per-CTA offsets without atomics.
constexpr int SG_BIN_THREADS = 256;

__global__ void __launch_bounds__(SG_BIN_THREADS) sg_classify_kernel(const int *__restrict__ size, int m, int chunk, SgBounds bnd,
                                                                     int *__restrict__ cls, int *__restrict__ block_count,
                                                                     int *__restrict__ class_count, int *__restrict__ max_size) {
    __shared__ int s_count[SG_NCLS];
    __shared__ int s_max;
    if (threadIdx.x < SG_NCLS) s_count[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    const int r0 = blockIdx.x * chunk, r1 = min(r0 + chunk, m);
    for (int i = r0 + threadIdx.x; i < r1; i += SG_BIN_THREADS) {
        const int u = size[i];
        int c = 0;
        if (u < 0) c = -u;  // explicit class (window / skip)
        else while (c < SG_NCLS - 1 && u > bnd.limit[c]) ++c;
        cls[i] = c;
        const unsigned peers = __match_any_sync(__activemask(), c);
        if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_count[c], __popc(peers));
        if (c == SG_GLOBAL_CLASS) atomicMax(&s_max, u);
    }
    __syncthreads();
    if (threadIdx.x < SG_NCLS) {
        const int v = s_count[threadIdx.x];
        block_count[blockIdx.x * SG_NCLS + threadIdx.x] = v;
        if (v) atomicAdd(&class_count[threadIdx.x], v);
    }
    if (threadIdx.x == 0 && s_max) atomicMax(max_size, s_max);
}

// block_count[b][c] -> first slot of CTA b's rows of class c: class start + counts of the CTAs before it
__global__ void sg_block_offsets_kernel(int *__restrict__ block_count, int nblocks, const int *__restrict__ class_start) {
    const int c = threadIdx.x;
    if (c >= SG_NCLS) return;
    int acc = class_start[c];
    for (int b = 0; b < nblocks; ++b) {
        const int v = block_count[b * SG_NCLS + c];
        block_count[b * SG_NCLS + c] = acc;
        acc += v;
    }
}

__global__ void __launch_bounds__(SG_BIN_THREADS) sg_bin_kernel(const int *__restrict__ cls, int m, int chunk,
                                                                const int *__restrict__ block_base, int *__restrict__ perm) {
    constexpr int NW = SG_BIN_THREADS / 32;
    __shared__ int s_cursor[SG_NCLS];
    __shared__ int s_wcount[NW][SG_NCLS];
    if (threadIdx.x < SG_NCLS) s_cursor[threadIdx.x] = block_base[blockIdx.x * SG_NCLS + threadIdx.x];
    const int r0 = blockIdx.x * chunk, r1 = min(r0 + chunk, m);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int base = r0; base < r1; base += SG_BIN_THREADS) {
        if (threadIdx.x < NW * SG_NCLS) (&s_wcount[0][0])[threadIdx.x] = 0;
        __syncthreads();
        const int i = base + threadIdx.x;
        int c = -1, rank = 0;
        if (i < r1) {
            c = cls[i];
            const unsigned peers = __match_any_sync(__activemask(), c);
            rank = __popc(peers & ((1u << lane) - 1u));
            if (lane == __ffs(peers) - 1) s_wcount[w][c] = __popc(peers);
        }
        __syncthreads();
        if (c >= 0) {
            int off = s_cursor[c];
            for (int v = 0; v < w; ++v) off += s_wcount[v][c];
            perm[off + rank] = i;
        }
        __syncthreads();
        if (threadIdx.x < SG_NCLS) {
            int t = 0;
            for (int v = 0; v < NW; ++v) t += s_wcount[v][threadIdx.x];
            s_cursor[threadIdx.x] += t;
        }
        __syncthreads();  // the counts are cleared at the top of the next round
    }
}

struct SgWin {
    const int *lo, *span;         // per row of A: all candidate columns lie in [lo, lo + span)
    int *saved;                   // per row: -1 not a window row | >= 0 offset of the finished row in tcol / tval | -2 deferred
    int *tcol; double *tval;      // scratch rows of the count pass
    unsigned long long *cursor; unsigned long long cap;
};

struct SgArgs {
    SgMat a, b;
    const int *perm; int count;   // rows of this class
    int hmask;                    // table size - 1 for this class
    int *row_nnz;                 // pass 1 output
    const int *c_rp; int *c_col; double *c_val;  // pass 2 output
    SgEpilogue ep;
    int *g_table; int *g_list;    // global-table class scratch (per CTA slices)
    int slots;                    // CTA class: a slot table fits next to the key table (2h ints + stage <= shared memory)
    SgWin win;                    // window class
};

// ---- tiny rows (ub <= SG_TINY): one THREAD per row.  A warp-per-row walk of a 7-product row is a
// chain of dependent loads with one lane busy; here 32 rows are in flight per warp.  The sorted key
// list lives in local memory; products are inserted / accumulated in ascending k, so pattern and
// values are the same as in the table kernels.
template <bool FILL>
__global__ void __launch_bounds__(128) sg_scalar_kernel(SgArgs s) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= s.count) return;
    const int i = s.perm[idx];
    int keys[SG_TINY];
    double vals[FILL ? SG_TINY : 1];
    int n = 0;
    const int a0 = s.a.rp[i], a1 = s.a.rp[i + 1];
    for (int q = a0; q < a1; ++q) {
        const int k = s.a.col[q];
        const double av = FILL ? s.a.val[q] : 0.0;
        const int b1 = s.b.rp[k + 1];
        for (int p = s.b.rp[k]; p < b1; ++p) {
            const int j = s.b.col[p];
            int pos = 0;
            while (pos < n && keys[pos] < j) ++pos;
            if (pos < n && keys[pos] == j) {
                if (FILL) vals[pos] = vals[pos] + av * s.b.val[p];
            } else {
                for (int t = n; t > pos; --t) { keys[t] = keys[t - 1]; if (FILL) vals[t] = vals[t - 1]; }
                keys[pos] = j;
                if (FILL) vals[pos] = 0.0 + av * s.b.val[p];
                ++n;
            }
        }
    }
    if (!FILL) { s.row_nnz[i] = n; return; }
    const int base = s.c_rp[i];
    if (!s.ep.enabled) {
        for (int t = 0; t < n; ++t) { s.c_col[base + t] = keys[t]; s.c_val[base + t] = vals[t]; }
        return;
    }
    double dv = 0.0; bool found = false;
    for (int q = a0; q < a1; ++q) if (s.a.col[q] == i) { dv = s.a.val[q]; found = true; break; }
    if (s.ep.enabled == 1 && (!found || !(dv > 1e-6))) atomicMax(s.ep.error_flag, 1);
    const double scalar = s.ep.omega * (1.0 / dv);
    const int p0 = s.ep.p.rp[i], p1 = s.ep.p.rp[i + 1];
    int matched = 0, pp = p0;
    for (int t = 0; t < n; ++t) {
        double v = sg_epi_scale(s.ep.enabled, vals[t], scalar);
        while (pp < p1 && s.ep.p.col[pp] < keys[t]) ++pp;
        if (pp < p1 && s.ep.p.col[pp] == keys[t]) { v = v + s.ep.p.val[pp]; ++matched; }
        s.c_col[base + t] = keys[t]; s.c_val[base + t] = v;
    }
    if (matched != p1 - p0) atomicMax(s.ep.error_flag, 2);
}

// one warp per row; dynamic shared memory: per warp (hmask+1) ints (+ half as many for the list)
template <bool FILL>
__global__ void __launch_bounds__(SG_WARPS * 32) sg_warp_kernel(SgArgs s) {
    extern __shared__ __align__(16) int sg_smem[];
    __shared__ int s_cnt[SG_WARPS];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = s.hmask + 1;
    const int per_warp = FILL ? 2 * h : h;  // fill: key table / accumulators (h) + sorted list (h/2) + slot table (h x u16)
    int *table = sg_smem + w * per_warp;
    const int idx = blockIdx.x * SG_WARPS + w;
    if (idx >= s.count) return;
    const int i = s.perm[idx];
    if (FILL) sg_row_fill<32>(i, lane, s.a, s.b, table, table + h, s.hmask, &s_cnt[w], s.c_rp, s.c_col, s.c_val, s.ep, nullptr,
                              reinterpret_cast<unsigned short *>(table + h + (h >> 1)));
    else sg_row_count<32>(i, lane, s.a, s.b, table, s.hmask, &s_cnt[w], s.row_nnz);
}

template <bool FILL>
__global__ void __launch_bounds__(256) sg_cta_kernel(SgArgs s) {
    extern __shared__ __align__(16) int sg_smem[];
    __shared__ int s_cnt;
    const int h = s.hmask + 1;
    const int i = s.perm[blockIdx.x];
    if (FILL) {
        SgStage *stage = reinterpret_cast<SgStage *>(sg_smem + (s.slots ? 2 * h : h + (h >> 1)));  // 8-byte aligned: h is a multiple of 4
        sg_row_fill<256>(i, threadIdx.x, s.a, s.b, sg_smem, sg_smem + h, s.hmask, &s_cnt, s.c_rp, s.c_col, s.c_val, s.ep, stage,
                         s.slots ? reinterpret_cast<unsigned short *>(sg_smem + h + (h >> 1)) : nullptr);
    } else sg_row_count<256>(i, threadIdx.x, s.a, s.b, sg_smem, s.hmask, &s_cnt, s.row_nnz);
}

// Pass 2 for CTA-class rows when B is narrow (ncols(B) <= SG_DENSE_MAX): one f64 accumulator per
// column of B in shared memory, so a product costs a load and a read-modify-write -- no hash table, no
// sort, no slot search (the binary search was 45 % of the instructions of the table kernel on the
// coarse-level products, ncu source view).  The structural pattern is recorded as one byte per
// column and compacted in column order at the end: sorted output, explicit zeros kept.
constexpr int SG_DENSE_MAX = 8192;

__global__ void __launch_bounds__(256) sg_dense_kernel(SgArgs s, int ncols_b) {
    extern __shared__ __align__(16) int sg_smem[];
    __shared__ int s_warp_tot[8];
    __shared__ int s_matched;
    double *acc = reinterpret_cast<double *>(sg_smem);
    SgStage *stage = reinterpret_cast<SgStage *>(acc + ncols_b);
    unsigned char *touched = reinterpret_cast<unsigned char *>(stage + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int i = s.perm[blockIdx.x];
    for (int t = tid; t < ncols_b; t += 256) { acc[t] = 0.0; touched[t] = 0; }
    if (tid == 0) s_matched = 0;
    __syncthreads();
    const int a0 = s.a.rp[i], a1 = s.a.rp[i + 1];
    sg_accumulate_staged<true>(tid, s.a, s.b, a0, a1, nullptr, 0, acc, *stage, touched);
    __syncthreads();
    // ordered compaction: thread t owns the columns [t*per, (t+1)*per)
    const int per = (ncols_b + 255) / 256;
    const int j0 = min(tid * per, ncols_b), j1 = min(j0 + per, ncols_b);
    int mine = 0;
    for (int j = j0; j < j1; ++j) mine += touched[j];
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    int before = incl - mine;
    for (int w = 0; w < warp; ++w) before += s_warp_tot[w];
    int pos = s.c_rp[i] + before;
    if (!s.ep.enabled) {
        for (int j = j0; j < j1; ++j)
            if (touched[j]) { s.c_col[pos] = j; s.c_val[pos] = acc[j]; ++pos; }
        return;
    }
    // smooth_interpolation epilogue (see sg_row_fill; linear diagonal search for the same reason)
    __shared__ int s_diag;
    if (tid == 0) s_diag = -1;
    __syncthreads();
    for (int q = a0 + tid; q < a1; q += 256) if (s.a.col[q] == i) s_diag = q;
    __syncthreads();
    const bool found = s_diag >= 0;
    const double dv = found ? s.a.val[s_diag] : 0.0;
    if (tid == 0 && s.ep.enabled == 1 && (!found || !(dv > 1e-6))) atomicMax(s.ep.error_flag, 1);
    const double scalar = s.ep.omega * (1.0 / dv);
    const int p0 = s.ep.p.rp[i], p1 = s.ep.p.rp[i + 1];
    int matched = 0;
    for (int j = j0; j < j1; ++j) {
        if (!touched[j]) continue;
        double v = sg_epi_scale(s.ep.enabled, acc[j], scalar);
        int lo = p0, hi = p1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (s.ep.p.col[mid] < j) lo = mid + 1; else hi = mid; }
        if (lo < p1 && s.ep.p.col[lo] == j) { v = v + s.ep.p.val[lo]; ++matched; }
        s.c_col[pos] = j; s.c_val[pos] = v; ++pos;
    }
    if (matched) atomicAdd(&s_matched, matched);
    __syncthreads();
    if (tid == 0 && s_matched != p1 - p0) atomicMax(s.ep.error_flag, 2);
}

// ---- window rows: one WARP per row, one accumulator per column of the row's window ---------------------------
// These products are bound by instruction issue, not by memory (ncu, r2: a CTA-per-row version of this kernel that gave every
// warp a range of the window's columns and walked the k's of a staged batch over 8 short pieces executed 7.8 warp
// instructions per product, 62 % issue-slot utilisation, 25 ms for the 2.3 G products of level 3's A*P).  So the walk is
// organised for few instructions per product instead:
//   * a warp owns the whole row: the entries of one B_k are taken 128 at a time (four independent read-modify-writes per lane,
//     all of one k: distinct columns, no conflicts), k after k in ascending order with nothing but __syncwarp in between -- no
//     CTA barriers, no per-piece bookkeeping, no search;
//   * the descriptors (start, length, a_ik) of 32 k's at a time live in registers (one per lane; the next 32 are fetched while
//     the current ones are processed), a warp scan of the lengths cuts them into chunks of <= WR_SE entries (a longer B_k is cut
//     into pieces that run one after the other: the order per output entry stays ascending k);
//   * the next chunk is brought into the warp's staging buffer with cp.async while the current one is accumulated.
// SAVE (count pass): the finished row goes to a bump-allocated scratch buffer and its length to row_nnz; the fill pass copies
// it to its place once row_ptr is known (sg_copy_saved_kernel).  Rows that did not fit are recomputed by the fill pass
// (!SAVE writes straight into C).
constexpr unsigned WR_FULL = 0xffffffffu;
constexpr size_t wn_smem_bytes(int span_max, int se) { return (size_t)9 * span_max + (size_t)24 * se; }

struct WrDesc { int b0, len, end; double av; };  // lane l: B-row of the l-th k of the group; end = inclusive scan of len

__device__ __forceinline__ WrDesc wr_load_group(const SgMat &a, const SgMat &b, int q, int a1, int lane) {
    WrDesc d{0, 0, 0, 0.0};
    if (q + lane < a1) {
        const int k = a.col[q + lane];
        d.av = a.val[q + lane];
        d.b0 = b.rp[k];
        d.len = b.rp[k + 1] - d.b0;
    }
    return d;
}
__device__ __forceinline__ void wr_scan(WrDesc &d, int lane) {
    int incl = d.len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(WR_FULL, incl, o); if (lane >= o) incl += v; }
    d.end = incl;
}

// segments ks .. ke-1 of a group; the first contributes first_cnt entries starting at poff, the others are whole;
// base = position of the chunk's first entry in the group's scan space
struct WrChunk { int ks, ke, poff, first_cnt, base; };

// next chunk of the group starting at (ks, poff); advances (ks, poff) past it
__device__ __forceinline__ WrChunk wr_next_chunk(const WrDesc &d, int lane, int &ks, int &poff, int se) {
    const int len_ks = __shfl_sync(WR_FULL, d.len, ks), end_ks = __shfl_sync(WR_FULL, d.end, ks);
    const int rem = len_ks - poff;
    WrChunk c;
    c.ks = ks; c.poff = poff; c.base = end_ks - rem;
    if (rem > se) {  // a piece of one long B_k
        c.ke = ks + 1; c.first_cnt = se; poff += se;
    } else {            // as many whole segments as fit (the scan is monotone: the lanes that fit are a prefix of ks..31)
        const unsigned fits = __ballot_sync(WR_FULL, lane >= ks && d.end - c.base <= se);
        c.ke = ks + __popc(fits); c.first_cnt = rem; ks = c.ke; poff = 0;
    }
    return c;
}

__device__ __forceinline__ void wr_issue(const SgMat &b, const WrDesc &d, const WrChunk &c, int lane, int *dc, double *dv) {
    for (int l = c.ks; l < c.ke; ++l) {
        const int b0 = __shfl_sync(WR_FULL, d.b0, l), len = __shfl_sync(WR_FULL, d.len, l), end = __shfl_sync(WR_FULL, d.end, l);
        const bool first = l == c.ks;
        const int src = b0 + (first ? c.poff : 0), cnt = first ? c.first_cnt : len, pos = first ? 0 : end - len - c.base;
        for (int e = lane; e < cnt; e += 32) {
            __pipeline_memcpy_async(dc + pos + e, b.col + src + e, 4);
            __pipeline_memcpy_async(dv + pos + e, b.val + src + e, 8);
        }
    }
}

__device__ __forceinline__ void wr_consume(const WrDesc &d, const WrChunk &c, int lane, const int *dc, const double *dv, double *acc,
                                           unsigned char *touched, int jmin) {
    for (int l = c.ks; l < c.ke; ++l) {
        const double av = __shfl_sync(WR_FULL, d.av, l);
        const int len = __shfl_sync(WR_FULL, d.len, l), end = __shfl_sync(WR_FULL, d.end, l);
        const bool first = l == c.ks;
        const int cnt = first ? c.first_cnt : len, pos = first ? 0 : end - len - c.base;
        const int *pc = dc + pos;
        const double *pv = dv + pos;
        for (int e = lane; e < cnt; e += 128) {
            int cc[4];
            double vv[4], aa[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool ok = e + 32 * u < cnt;
                cc[u] = ok ? pc[e + 32 * u] - jmin : -1;
                vv[u] = ok ? pv[e + 32 * u] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) aa[u] = cc[u] >= 0 ? acc[cc[u]] : 0.0;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (cc[u] >= 0) { acc[cc[u]] = aa[u] + av * vv[u]; touched[cc[u]] = 1; }
        }
        __syncwarp();  // the next k may reach the same column from another lane
    }
}

template <bool SAVE>
__global__ void __launch_bounds__(32) sg_window_kernel(SgArgs s, int span_max, int se) {
    extern __shared__ __align__(16) unsigned char wn_smem[];
    double *acc = reinterpret_cast<double *>(wn_smem);                    // span_max
    double *sval = acc + span_max;                                     // 2 x se
    int *scol = reinterpret_cast<int *>(sval + 2 * se);                // 2 x se
    unsigned *touched_w = reinterpret_cast<unsigned *>(scol + 2 * se);  // span_max bytes
    unsigned char *touched = reinterpret_cast<unsigned char *>(touched_w);
    const int lane = threadIdx.x;
    const int i = s.perm[blockIdx.x];
    const int jmin = s.win.lo[i], span = s.win.span[i];
    const int span4 = (span + 3) >> 2;
    for (int t = lane; t < span; t += 32) acc[t] = 0.0;
    for (int t = lane; t < span4; t += 32) touched_w[t] = 0u;
    __syncwarp();
    const int a0 = s.a.rp[i], a1 = s.a.rp[i + 1];
    if (a0 < a1) {
        int qn = a0;  // first k of the group held in dn
        WrDesc dn = wr_load_group(s.a, s.b, qn, a1, lane), dcur = dn;
        wr_scan(dn, lane);
        int ks = 0, poff = 0, buf = 0;
        WrChunk cur = wr_next_chunk(dn, lane, ks, poff, se);
        wr_issue(s.b, dn, cur, lane, scol, sval);
        __pipeline_commit();
        bool cur_in_next = true, n_scanned = true;
        while (true) {
            if (cur_in_next) {  // the group `cur` was cut from is the one consumed now; fetch the descriptors of the one after it
                dcur = dn; qn += 32;
                if (qn < a1) { dn = wr_load_group(s.a, s.b, qn, a1, lane); n_scanned = false; }
                cur_in_next = false;
            }
            WrChunk nxt = cur;
            bool have = false, nxt_in_next = false;
            if (ks < 32) {
                nxt = wr_next_chunk(dcur, lane, ks, poff, se); have = true;
            } else if (qn < a1) {
                if (!n_scanned) { wr_scan(dn, lane); n_scanned = true; }
                ks = 0; poff = 0;
                nxt = wr_next_chunk(dn, lane, ks, poff, se); have = true; nxt_in_next = true;
            }
            if (have) {
                if (nxt_in_next) wr_issue(s.b, dn, nxt, lane, scol + (buf ^ 1) * se, sval + (buf ^ 1) * se);
                else wr_issue(s.b, dcur, nxt, lane, scol + (buf ^ 1) * se, sval + (buf ^ 1) * se);
            }
            __pipeline_commit();
            __pipeline_wait_prior(1);  // everything but the chunk just issued has landed
            __syncwarp();
            wr_consume(dcur, cur, lane, scol + buf * se, sval + buf * se, acc, touched, jmin);
            if (!have) break;
            cur = nxt; buf ^= 1; cur_in_next = nxt_in_next;
        }
    }
    __syncwarp();
    // length of the row, then its entries in column order
    int n = 0;
    for (int t = lane; t < span4; t += 32) n += __popc(touched_w[t] & 0x01010101u);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(WR_FULL, n, o);
    int *o_col; double *o_val;
    if (SAVE) {
        long long base = 0;
        if (lane == 0) {
            s.row_nnz[i] = n;
            const unsigned long long off = atomicAdd(s.win.cursor, (unsigned long long)n);
            const bool fits = off + (unsigned long long)n <= s.win.cap;
            s.win.saved[i] = fits ? (int)off : -2;
            base = fits ? (long long)off : -1;
        }
        base = __shfl_sync(WR_FULL, base, 0);
        if (base < 0) return;
        o_col = s.win.tcol + base; o_val = s.win.tval + base;
    } else {
        o_col = s.c_col + s.c_rp[i]; o_val = s.c_val + s.c_rp[i];
    }
    // 128 columns per step: lane l holds the four flags of columns 4 (w0 + l) .. + 3 (one word of `touched`)
    int run = 0;
    const unsigned lt = (1u << lane) - 1u;
    for (int w0 = 0; w0 < span4; w0 += 32) {
        const int wi = w0 + lane;
        const unsigned w = wi < span4 ? touched_w[wi] : 0u;
        const unsigned m0 = __ballot_sync(WR_FULL, w & 0x1u), m1 = __ballot_sync(WR_FULL, w & 0x100u),
                       m2 = __ballot_sync(WR_FULL, w & 0x10000u), m3 = __ballot_sync(WR_FULL, w & 0x1000000u);
        int pos = run + __popc(m0 & lt) + __popc(m1 & lt) + __popc(m2 & lt) + __popc(m3 & lt);
        const int j = 4 * wi;
        if (w & 0x1u) { o_col[pos] = jmin + j; o_val[pos] = acc[j]; ++pos; }
        if (w & 0x100u) { o_col[pos] = jmin + j + 1; o_val[pos] = acc[j + 1]; ++pos; }
        if (w & 0x10000u) { o_col[pos] = jmin + j + 2; o_val[pos] = acc[j + 2]; ++pos; }
        if (w & 0x1000000u) { o_col[pos] = jmin + j + 3; o_val[pos] = acc[j + 3]; }
        run += __popc(m0) + __popc(m1) + __popc(m2) + __popc(m3);
    }
}

// fill pass of the window rows the count pass finished: scratch row -> its place in C
__global__ void __launch_bounds__(128) sg_copy_saved_kernel(const int *__restrict__ rows, const int *__restrict__ saved,
                                                            const int *__restrict__ tcol, const double *__restrict__ tval,
                                                            const int *__restrict__ c_rp, int *__restrict__ c_col, double *__restrict__ c_val) {
    const int i = rows[blockIdx.x];
    const int off = saved[i];
    if (off < 0) return;
    const int base = c_rp[i], n = c_rp[i + 1] - base;
    for (int t = threadIdx.x; t < n; t += 128) { c_col[base + t] = tcol[off + t]; c_val[base + t] = tval[off + t]; }
}

// Pass 1 for rows with many candidates when B is not too wide: distinct columns counted with a
// bitmap of ncols(B) bits in shared memory (clearing ncols/8 bytes beats clearing a hash table of
// 4 * 1.33 * ub bytes for the dense rows of coarse-level products).
__global__ void __launch_bounds__(256) sg_bitmap_count_kernel(SgArgs s, int words) {
    extern __shared__ __align__(16) int sg_smem[];
    __shared__ int s_cnt;
    unsigned *bm = reinterpret_cast<unsigned *>(sg_smem);
    const int i = s.perm[blockIdx.x];
    for (int t = threadIdx.x; t < words; t += 256) bm[t] = 0u;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const int a0 = s.a.rp[i], a1 = s.a.rp[i + 1];
    for (int q = a0 + (threadIdx.x >> 3); q < a1; q += 32) {
        const int k = s.a.col[q];
        const int b1 = s.b.rp[k + 1];
        for (int p = s.b.rp[k] + (threadIdx.x & 7); p < b1; p += 8) {
            const int j = s.b.col[p];
            atomicOr(&bm[j >> 5], 1u << (j & 31));
        }
    }
    __syncthreads();
    int c = 0;
    for (int t = threadIdx.x; t < words; t += 256) c += __popc(bm[t]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) s.row_nnz[i] = s_cnt;
}

template <bool FILL>
__global__ void __launch_bounds__(256) sg_global_kernel(SgArgs s) {
    __shared__ int s_cnt;
    const size_t h = (size_t)s.hmask + 1;
    int *table = s.g_table + (size_t)blockIdx.x * h;
    int *list = s.g_list + (size_t)blockIdx.x * (h >> 1);
    for (int idx = blockIdx.x; idx < s.count; idx += gridDim.x) {
        const int i = s.perm[idx];
        if (FILL) sg_row_fill<256>(i, threadIdx.x, s.a, s.b, table, list, s.hmask, &s_cnt, s.c_rp, s.c_col, s.c_val, s.ep);
        else sg_row_count<256>(i, threadIdx.x, s.a, s.b, table, s.hmask, &s_cnt, s.row_nnz);
        __syncthreads();
    }
}

// Class plan.  Pass 1 sizes tables from the upper bound ub (keys only, load factor <= 0.75);
// pass 2 from the exact row length (load factor <= 0.5; 6 bytes of shared memory per table slot).
//   kind 0: warp per row (SG_WARPS rows per CTA), 1: CTA per row, 2: CTA per row, global tables
struct SgClass { int limit, h, kind; };
// Only the last class is open-ended (it alone tracks the maximum row size for its table); unused
// slots repeat the previous limit so that nothing falls into them.
//   kind 3: thread per row, 4: window rows (sg_window_kernel), 5: nothing to do (fill pass: rows the count pass finished)
// Classes SG_SKIP_CLASS and WN_CLASS are only entered explicitly (negative sizes); SG_GLOBAL_CLASS is the open-ended one and
// alone tracks the maximum row size for its table.
static const SgClass SG_PASS1[SG_NCLS] = {{SG_TINY, 0, 3},   {192, 256, 0}, {768, 1024, 0},     {3072, 4096, 0},
                                          {12288, 16384, 1}, {12288, 0, 5}, {0x7fffffff, 0, 2}, {0x7fffffff, 0, 4}};
static const SgClass SG_PASS2[SG_NCLS] = {{0, 0, 3},         {128, 256, 0}, {512, 1024, 0},     {2048, 4096, 1},
                                          {16384, 32768, 1}, {16384, 0, 5}, {0x7fffffff, 0, 2}, {0x7fffffff, 0, 4}};

// host side of the window class for one product
struct SgWinHost {
    bool on = false;
    int *blk = nullptr;          // lo | span | saved (m + 2 each) | bfirst | blast (nrows(B) + 2 each)
    int *lo = nullptr, *span = nullptr, *saved = nullptr;
    unsigned long long *bound = nullptr, *cursor = nullptr;  // in the counters block
    int span_max = 0;
    int *tcol = nullptr; double *tval = nullptr; unsigned long long cap = 0;
    int rows_off = 0, rows = 0;  // slice of `perm` holding the window rows of the count pass
};
static unsigned long long sg_win_cap() {  // entries of the scratch buffer at most (12 bytes each)
    static const unsigned long long cap = [] { const char *e = getenv("FAMG_SG_WIN_CAP"); return e ? strtoull(e, nullptr, 10) : (32ull << 20); }();
    return cap;
}

template <bool FILL>
static famg_status sg_run_pass(famg_ctx *ctx, SgArgs base, const int *d_size, int m, int ncols_b, int *d_cls, int *d_perm,
                               int *d_counters, int **scratch, SgWinHost *wh) {
    const SgClass *plan = FILL ? SG_PASS2 : SG_PASS1;
    SgBounds bnd;
    for (int c = 0; c < SG_NCLS; ++c) bnd.limit[c] = plan[c].limit;
    int h_counters[32] = {0};  // [0..7] class counts, [8] largest size of the global class, [12] error flag, [16..] window class
    cudaMemsetAsync(d_counters, 0, sizeof(int) * 16, ctx->stream);
    // contiguous chunks of rows per CTA (multiples of the CTA width so warps stay full)
    const int nblocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(m, SG_BIN_THREADS * 4), 8 * (int64_t)ctx->num_sms));
    const int chunk = (int)(ceil_div(ceil_div(m, nblocks), SG_BIN_THREADS) * SG_BIN_THREADS);
    int *d_block = nullptr;
    const size_t block_bytes = sizeof(int) * (size_t)nblocks * SG_NCLS;
    FAMG_TRY(pool_alloc(ctx, block_bytes, (void **)&d_block));
    sg_classify_kernel<<<nblocks, SG_BIN_THREADS, 0, ctx->stream>>>(d_size, m, chunk, bnd, d_cls, d_block, d_counters, d_counters + 8);
    count_launch(ctx);
    cudaError_t ce = cudaMemcpyAsync(h_counters, d_counters, sizeof(int) * 32, cudaMemcpyDeviceToHost, ctx->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
    int h_cursor[SG_NCLS], acc = 0;
    for (int c = 0; c < SG_NCLS; ++c) { h_cursor[c] = acc; acc += h_counters[c]; }
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_counters, h_cursor, sizeof(int) * SG_NCLS, cudaMemcpyHostToDevice, ctx->stream);
    if (ce == cudaSuccess) {
        sg_block_offsets_kernel<<<1, 32, 0, ctx->stream>>>(d_block, nblocks, d_counters);
        sg_bin_kernel<<<nblocks, SG_BIN_THREADS, 0, ctx->stream>>>(d_cls, m, chunk, d_block, d_perm);
        count_launch(ctx, 2);
        ce = cudaStreamSynchronize(ctx->stream);  // h_cursor is a stack buffer
    }
    pool_free(ctx, d_block, block_bytes);
    if (ce != cudaSuccess) FAMG_FAIL(FAMG_ERR_CUDA, "spgemm binning failed: %s", cudaGetErrorString(ce));
    const int max_size = h_counters[8];
    int off = 0;
    // merge the open-ended global classes into one launch
    int global_count = 0, global_off = -1;
    for (int c = 0; c < SG_NCLS; ++c) {
        const int cnt = h_counters[c];
        SgArgs s = base;
        s.perm = d_perm + off; s.count = cnt;
        const int my_off = off;
        off += cnt;
        if (cnt == 0 || plan[c].kind == 5) continue;
        if (plan[c].kind == 4) {
            if (!wh || !wh->on) FAMG_FAIL(FAMG_ERR_CUDA, "spgemm: window rows without a window plan");
            if (!FILL) {
                unsigned long long bound = 0;
                memcpy(&bound, h_counters + 16, sizeof(bound));
                wh->span_max = (h_counters[18] + 15) & ~15;
                wh->cap = std::max<unsigned long long>(std::min(bound, sg_win_cap()), 1);
                FAMG_TRY(pool_alloc(ctx, sizeof(int) * (size_t)wh->cap, (void **)&wh->tcol));
                FAMG_TRY(pool_alloc(ctx, sizeof(double) * (size_t)wh->cap, (void **)&wh->tval));
                wh->rows_off = my_off; wh->rows = cnt;
            }
            s.win = SgWin{wh->lo, wh->span, wh->saved, wh->tcol, wh->tval, wh->cursor, wh->cap};
            // FAMG_SG_WIN_SE: other chunk sizes for measurements (1024-entry chunks where they cost no resident row changed nothing:
            // the rows are bound by the instruction latency of their one warp, not by the prefetch distance)
            static const int se_forced = [] { const char *e = getenv("FAMG_SG_WIN_SE"); return e ? atoi(e) : 0; }();
            const int se = se_forced > 0 ? se_forced : WR_SE;
            const size_t smem = wn_smem_bytes(wh->span_max, se);
            CUDA_TRY(cudaFuncSetAttribute(sg_window_kernel<!FILL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            sg_window_kernel<!FILL><<<(unsigned)cnt, 32, smem, ctx->stream>>>(s, wh->span_max, se);
            count_launch(ctx);
            KERNEL_CHECK();
            continue;
        }
        const int bm_words = (ncols_b + 31) / 32;
        if (!FILL && plan[c].kind >= 1 && (size_t)bm_words * 4 <= 160 * 1024) {
            // big rows of the count pass: bitmap over the columns of B
            const size_t smem = (size_t)bm_words * 4;
            if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(sg_bitmap_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            sg_bitmap_count_kernel<<<(unsigned)cnt, 256, smem, ctx->stream>>>(s, bm_words);
            count_launch(ctx);
            KERNEL_CHECK();
            continue;
        }
        if (FILL && plan[c].kind >= 1 && ncols_b <= SG_DENSE_MAX) {
            const size_t smem = (size_t)ncols_b * 8 + sizeof(SgStage) + (((size_t)ncols_b + 15) & ~(size_t)15);
            if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(sg_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
            sg_dense_kernel<<<(unsigned)cnt, 256, smem, ctx->stream>>>(s, ncols_b);
            count_launch(ctx);
            KERNEL_CHECK();
            continue;
        }
        if (plan[c].kind == 2) { if (global_off < 0) global_off = my_off; global_count += cnt; continue; }
        if (plan[c].kind == 3) {  // tiny rows: one thread per row
            sg_scalar_kernel<FILL><<<(unsigned)ceil_div(cnt, 128), 128, 0, ctx->stream>>>(s);
            count_launch(ctx);
            KERNEL_CHECK();
            continue;
        }
        s.hmask = plan[c].h - 1;
        // fill: key table / accumulators (h ints) + sorted list (h/2) + slot table (h x u16) when it fits (the 32768-slot class does not)
        s.slots = FILL && (plan[c].kind == 0 || (size_t)2 * plan[c].h * sizeof(int) + sizeof(SgStage) <= 200 * 1024);
        const size_t per_row = (size_t)(FILL ? (s.slots ? 2 * plan[c].h : plan[c].h + plan[c].h / 2) : plan[c].h) * sizeof(int);
        if (plan[c].kind == 0) {
            const size_t smem = per_row * SG_WARPS;
            if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(sg_warp_kernel<FILL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            sg_warp_kernel<FILL><<<(unsigned)ceil_div(cnt, SG_WARPS), SG_WARPS * 32, smem, ctx->stream>>>(s);
        } else {
            const size_t cta_smem = per_row + (FILL ? sizeof(SgStage) : 0);
            if (cta_smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(sg_cta_kernel<FILL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));  // 227 KB opt-in limit minus the static s_cnt
            sg_cta_kernel<FILL><<<(unsigned)cnt, 256, cta_smem, ctx->stream>>>(s);
        }
        count_launch(ctx);
        KERNEL_CHECK();
    }
    if (global_count > 0) {
        if (max_size > (1 << 27)) FAMG_FAIL(FAMG_ERR_UNSUPPORTED, "spgemm: a row of the product has %d candidate entries", max_size);
        int h = 1;
        while (h < 2 * max_size) h <<= 1;
        const int grid = std::min(global_count, 2 * ctx->num_sms);
        if (*scratch) { pool_free(ctx, *scratch, 0); *scratch = nullptr; }
        FAMG_TRY(pool_alloc(ctx, sizeof(int) * (size_t)grid * (size_t)(h + h / 2), (void **)scratch));
        SgArgs s = base;
        s.perm = d_perm + global_off; s.count = global_count; s.hmask = h - 1;
        s.g_table = *scratch; s.g_list = *scratch + (size_t)grid * h;
        sg_global_kernel<FILL><<<grid, 256, 0, ctx->stream>>>(s);
        count_launch(ctx);
        KERNEL_CHECK();
    }
    return FAMG_OK;
}

famg_status spgemm_impl(const famg_csr *a, const famg_csr *b, const famg_csr *p_for_smoothing, double omega, famg_csr **out, int epi_kind) {
    *out = nullptr;
    if (a->ncols != b->nrows) FAMG_FAIL(FAMG_ERR_INVALID, "spgemm: inner dimensions differ (%lld vs %lld)", (long long)a->ncols, (long long)b->nrows);
    famg_ctx *ctx = a->ctx;
    const int m = (int)a->nrows;
    if (m == 0) {  // empty product: nothing to count, nothing to read back
        FAMG_TRY(csr_alloc(ctx, 0, b->ncols, 0, out));
        return csr_finalize_plan(*out);
    }
    SgMat A{a->row_ptr, a->col, a->val}, B{b->row_ptr, b->col, b->val};
    // one pooled scratch block: ub | size | cls | perm | row_nnz(+1) | rp(+1) | counters(32)
    const size_t words = (size_t)6 * (m + 2) + 32;
    int *blk = nullptr, *scratch = nullptr;
    famg_csr *c = nullptr;
    famg_status st;
    { PhaseTimer pt(ctx, "    spgemm: scratch"); st = pool_alloc(ctx, words * sizeof(int), (void **)&blk); }
    if (st != FAMG_OK) return st;
    int *ub = blk, *size = ub + (m + 2), *cls = size + (m + 2), *perm = cls + (m + 2), *row_nnz = perm + (m + 2),
        *rp = row_nnz + (m + 2), *counters = rp + (m + 2);
    SgWinHost wh;
    auto cleanup = [&]() {
        pool_free(ctx, blk, words * sizeof(int));
        if (scratch) pool_free(ctx, scratch, 0);
        if (wh.blk) pool_free(ctx, wh.blk, 0);
        if (wh.tcol) pool_free(ctx, wh.tcol, 0);
        if (wh.tval) pool_free(ctx, wh.tval, 0);
    };
#define SG_TRY(expr) do { st = (expr); if (st != FAMG_OK) { cleanup(); if (c) csr_release(c); return st; } } while (0)
    SgArgs base{};
    base.a = A; base.b = B; base.row_nnz = row_nnz;
    base.ep.enabled = 0;
    int total = 0;
    // window class (sg_window_kernel): worth its set-up only when rows with thousands of products are to be expected
    static const int win_min_ub = [] { const char *e = getenv("FAMG_SG_WIN_MIN_UB"); return e ? atoi(e) : 3072; }();
    static const bool win_enabled = [] { const char *e = getenv("FAMG_SG_WINDOW"); return !e || atoi(e) != 0; }();
    static const int win_min_seg = [] { const char *e = getenv("FAMG_SG_WIN_MIN_SEG"); return e ? atoi(e) : 64; }();
    const double est_ub = (double)a->nnz / (double)m * ((double)b->nnz / (double)std::max<int64_t>(b->nrows, 1));
    SgWinPlan wp{};
    if (win_enabled && !p_for_smoothing && est_ub * 4 >= win_min_ub && b->nrows > 0) {
        const int nb = (int)b->nrows;
        SG_TRY(pool_alloc(ctx, sizeof(int) * ((size_t)3 * (m + 2) + (size_t)2 * (nb + 2)), (void **)&wh.blk));
        wh.on = true;
        wh.lo = wh.blk; wh.span = wh.lo + (m + 2); wh.saved = wh.span + (m + 2);
        int *bfirst = wh.saved + (m + 2), *blast = bfirst + (nb + 2);
        wh.bound = reinterpret_cast<unsigned long long *>(counters + 16);
        wh.cursor = reinterpret_cast<unsigned long long *>(counters + 20);
        cudaMemsetAsync(counters + 16, 0, sizeof(int) * 16, ctx->stream);
        cudaMemsetAsync(wh.saved, 0xff, sizeof(int) * (size_t)(m + 2), ctx->stream);
        sg_bspan_kernel<<<(unsigned)ceil_div((int64_t)nb * 8, 256), 256, 0, ctx->stream>>>(B, nb, bfirst, blast, counters + 19);
        count_launch(ctx);
        wp = SgWinPlan{bfirst, blast, counters + 19, wh.lo, wh.span, wh.bound, counters + 18, win_min_ub, win_min_seg};
    }
    if (m > 0) {
        PhaseTimer pt(ctx, "    spgemm: bounds + count pass");
        sg_ub_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, ctx->stream>>>(A, B, m, (int)b->ncols, ub, size, wp);
        count_launch(ctx);
        SG_TRY((sg_run_pass<false>(ctx, base, size, m, (int)b->ncols, cls, perm, counters, &scratch, &wh)));
    }
    PhaseTimer pt_alloc(ctx, "    spgemm: scan + allocate C");
    SG_TRY(exclusive_scan_i32(ctx, row_nnz, rp, m));
    {
        cudaError_t e = cudaMemcpy(&total, rp + m, sizeof(int), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_error("spgemm: %s", cudaGetErrorString(e)); SG_TRY(FAMG_ERR_CUDA); }
        if (total < 0) { set_error("spgemm: product has more than 2^31 non-zeros"); SG_TRY(FAMG_ERR_UNSUPPORTED); }
    }
    SG_TRY(csr_alloc(ctx, a->nrows, b->ncols, total, &c));
    cudaMemcpyAsync(c->row_ptr, rp, sizeof(int) * (m + 1), cudaMemcpyDeviceToDevice, ctx->stream);
    int *err_flag = counters + 12;
    base.c_rp = c->row_ptr; base.c_col = c->col; base.c_val = c->val;
    if (p_for_smoothing) {
        base.ep.enabled = epi_kind; base.ep.omega = omega;
        base.ep.p = SgMat{p_for_smoothing->row_ptr, p_for_smoothing->col, p_for_smoothing->val};
        base.ep.error_flag = err_flag;
    }
    pt_alloc.~PhaseTimer(); pt_alloc.on = false;
    if (m > 0) {
        PhaseTimer pt(ctx, "    spgemm: fill pass");
        if (wh.rows > 0) {  // rows the count pass finished (`perm` still holds the count pass's order)
            sg_copy_saved_kernel<<<(unsigned)wh.rows, 128, 0, ctx->stream>>>(perm + wh.rows_off, wh.saved, wh.tcol, wh.tval, c->row_ptr, c->col, c->val);
            count_launch(ctx);
        }
        sg_size2_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, ctx->stream>>>(row_nnz, ub, m, size, wh.on ? wh.saved : nullptr);
        count_launch(ctx);
        SG_TRY((sg_run_pass<true>(ctx, base, size, m, (int)b->ncols, cls, perm, counters, &scratch, &wh)));
    }
    int h_err = 0;
    if (p_for_smoothing) {
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
    famg_status st = pool_alloc(ctx, sizeof(int) * (size_t)(n + 1), (void **)&counts);
    if (st == FAMG_OK) st = pool_alloc(ctx, sizeof(int) * (size_t)(n + 1), (void **)&cursor);
    if (st == FAMG_OK) st = pool_alloc(ctx, sizeof(int) * (size_t)std::max(nnz, 1), (void **)&t_col);
    if (st == FAMG_OK) st = pool_alloc(ctx, sizeof(double) * (size_t)std::max(nnz, 1), (void **)&t_val);
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
    pool_free(ctx, counts, 0); pool_free(ctx, cursor, 0); pool_free(ctx, t_col, 0); pool_free(ctx, t_val, 0);  // stream-ordered
    if (st == FAMG_OK) st = csr_finalize_plan(t);
    if (st != FAMG_OK) { csr_release(t); return st; }
    *out = t;
    return FAMG_OK;
}

// ---------------------------------------------------------------- block Jacobi prolongator smoothing
// diagonal blocks of A (block_size x block_size, row-major per block, absent entries 0): one thread per row
__global__ void block_diag_extract_kernel(const int *__restrict__ rp, const int *__restrict__ col, const double *__restrict__ val, int n,
                                          int bs, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int start = (i / bs) * bs;
    double *o = out + (size_t)i * bs;
    for (int j = 0; j < bs; ++j) o[j] = 0.0;
    for (int q = rp[i]; q < rp[i + 1]; ++q) {
        const int j = col[q] - start;
        if (j >= 0 && j < bs) o[j] = val[q];
    }
}

// cyclic Jacobi eigen-decomposition of a small symmetric matrix (column-major, both triangles filled):
// on return `m` holds the eigenvalues on its diagonal and `u` the eigenvectors in its columns
static void host_sym_eig(int n, std::vector<double> &m, std::vector<double> &u) {
    u.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) u[(size_t)i + (size_t)i * n] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int p = 0; p < n; ++p) {
            diag += m[(size_t)p + (size_t)p * n] * m[(size_t)p + (size_t)p * n];
            for (int q = p + 1; q < n; ++q) off += m[(size_t)p + (size_t)q * n] * m[(size_t)p + (size_t)q * n];
        }
        if (off <= 1e-32 * diag || off == 0.0) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = m[(size_t)p + (size_t)q * n];
                if (apq == 0.0) continue;
                const double app = m[(size_t)p + (size_t)p * n], aqq = m[(size_t)q + (size_t)q * n];
                const double zeta = (aqq - app) / (2.0 * apq);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
                for (int r = 0; r < n; ++r) {  // columns p, q
                    const double mp = m[(size_t)r + (size_t)p * n], mq = m[(size_t)r + (size_t)q * n];
                    m[(size_t)r + (size_t)p * n] = c * mp - sn * mq;
                    m[(size_t)r + (size_t)q * n] = sn * mp + c * mq;
                }
                for (int r = 0; r < n; ++r) {  // rows p, q
                    const double mp = m[(size_t)p + (size_t)r * n], mq = m[(size_t)q + (size_t)r * n];
                    m[(size_t)p + (size_t)r * n] = c * mp - sn * mq;
                    m[(size_t)q + (size_t)r * n] = sn * mp + c * mq;
                }
                for (int r = 0; r < n; ++r) {
                    const double up = u[(size_t)r + (size_t)p * n], uq = u[(size_t)r + (size_t)q * n];
                    u[(size_t)r + (size_t)p * n] = c * up - sn * uq;
                    u[(size_t)r + (size_t)q * n] = sn * up + c * uq;
                }
            }
    }
}

// D^-1 of interpolation/mod.rs:963-1015: scale * (U diag(1/s) U^T) of every diagonal block (lower side),
// all block_size^2 entries kept (the reference pushes a triplet for each), as a device CSR
static famg_status block_diag_inverse(const famg_csr *a, int bs, double scale, famg_csr **out) {
    famg_ctx *ctx = a->ctx;
    const int64_t n = a->nrows, nb = n / bs;
    double *d_blocks = nullptr;
    const size_t bytes = sizeof(double) * (size_t)n * bs;
    FAMG_TRY(pool_alloc(ctx, bytes, (void **)&d_blocks));
    if (n > 0) {
        block_diag_extract_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, ctx->stream>>>(a->row_ptr, a->col, a->val, (int)n, bs, d_blocks);
        count_launch(ctx);
    }
    std::vector<double> blocks((size_t)n * bs);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(blocks.data(), d_blocks, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    pool_free(ctx, d_blocks, bytes);
    if (e != cudaSuccess) FAMG_FAIL(FAMG_ERR_CUDA, "block_jacobi: %s", cudaGetErrorString(e));
    std::vector<int> rp((size_t)n + 1), ci((size_t)n * bs);
    std::vector<double> cv((size_t)n * bs);
    double worst = 1.0;
    bool bad = false;
#pragma omp parallel
    {
        std::vector<double> m, u;
#pragma omp for schedule(static)
        for (int64_t b = 0; b < nb; ++b) {
            m.assign((size_t)bs * bs, 0.0);
            for (int i = 0; i < bs; ++i)
                for (int j = 0; j <= i; ++j) {  // Side::Lower
                    const double v = blocks[((size_t)b * bs + i) * bs + j];
                    m[(size_t)i + (size_t)j * bs] = v; m[(size_t)j + (size_t)i * bs] = v;
                }
            host_sym_eig(bs, m, u);
            for (int c = 0; c < bs; ++c) {
                const double ev = m[(size_t)c + (size_t)c * bs];
                if (!(ev > 1e-6)) {
#pragma omp critical
                    { bad = true; worst = std::min(worst, ev); }
                }
            }
            for (int i = 0; i < bs; ++i)
                for (int j = 0; j < bs; ++j) {
                    double sum = 0.0;
                    for (int c = 0; c < bs; ++c) sum += u[(size_t)i + (size_t)c * bs] * (1.0 / m[(size_t)c + (size_t)c * bs]) * u[(size_t)j + (size_t)c * bs];
                    const size_t row = (size_t)b * bs + i;
                    ci[row * bs + j] = (int)(b * bs + j);
                    cv[row * bs + j] = scale * sum;
                }
        }
    }
    if (bad) FAMG_FAIL(FAMG_ERR_NUMERIC, "block diagonal is nearly singular with eigval of: %.3e", worst);  // interpolation/mod.rs:994-998
    for (int64_t i = 0; i <= n; ++i) rp[(size_t)i] = (int)(i * bs);
    return csr_from_host_i32(ctx, n, n, rp.data(), ci.data(), cv.data(), out);
}

// block_jacobi (interpolation/mod.rs:963-1028): D^-1 = -0.66 (block diag of A)^-1; S = D^-1 (A P); S += P
static famg_status block_jacobi_impl(const famg_csr *a, int bs, const famg_csr *p, famg_csr **out) {
    famg_csr *dinv = nullptr, *ap = nullptr;
    famg_status st = block_diag_inverse(a, bs, -JACOBI_WEIGHT_BLOCK, &dinv);
    if (st == FAMG_OK) st = spgemm_impl(a, p, nullptr, 0.0, &ap);
    if (st == FAMG_OK) st = spgemm_impl(dinv, ap, p, 0.0, out, 2);
    if (dinv) csr_release(dinv);
    if (ap) csr_release(ap);
    return st;
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

famg_status famg_block_jacobi(const famg_csr *a, int64_t block_size, const famg_csr *p, famg_csr **out) {
    if (!a || !p || !out || block_size < 1) FAMG_FAIL(FAMG_ERR_INVALID, "bad argument");
    if (a->nrows != a->ncols || p->nrows != a->nrows || a->nrows % block_size != 0 || block_size > 64)
        FAMG_FAIL(FAMG_ERR_INVALID, "block_jacobi shape mismatch");
    CUDA_TRY(cudaSetDevice(a->ctx->device));
    *out = nullptr;
    return block_jacobi_impl(a, (int)block_size, p, out);
}

famg_status famg_smooth_p(const famg_csr *a, const famg_csr *m_inv, const famg_csr *p, famg_csr **out) {
    if (!a || !m_inv || !p || !out) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    if (a->nrows != a->ncols || m_inv->ncols != a->nrows || m_inv->nrows != p->nrows || p->nrows != a->nrows)
        FAMG_FAIL(FAMG_ERR_INVALID, "smooth_p shape mismatch");
    CUDA_TRY(cudaSetDevice(a->ctx->device));
    *out = nullptr;
    famg_csr *ap = nullptr;
    famg_status st = spgemm_impl(a, p, nullptr, 0.0, &ap);
    if (st == FAMG_OK) st = spgemm_impl(m_inv, ap, p, 0.0, out, 3);  // m_inv * (-(A P)) + P
    if (ap) csr_release(ap);
    return st;
}

famg_status famg_galerkin(const famg_csr *a, const famg_csr *p0, int smoothing_steps, double omega, famg_csr **p_out,
                          famg_csr **r_out, famg_csr **ac_out) {
    return famg_galerkin_block(a, p0, 1, smoothing_steps, omega, p_out, r_out, ac_out);
}

famg_status famg_galerkin_block(const famg_csr *a, const famg_csr *p0, int64_t block_size, int smoothing_steps, double omega,
                                famg_csr **p_out, famg_csr **r_out, famg_csr **ac_out) {
    if (!a || !p0 || !p_out || !r_out || !ac_out || smoothing_steps < 0 || block_size < 1 || block_size > 64)
        FAMG_FAIL(FAMG_ERR_INVALID, "bad argument");
    *p_out = *r_out = *ac_out = nullptr;
    if (a->nrows != a->ncols || p0->nrows != a->nrows || a->nrows % block_size != 0) FAMG_FAIL(FAMG_ERR_INVALID, "galerkin shape mismatch");
    CUDA_TRY(cudaSetDevice(a->ctx->device));
    famg_csr *p = const_cast<famg_csr *>(p0);
    p->refs.fetch_add(1);
    famg_status st = FAMG_OK;
    for (int s = 0; s < smoothing_steps && st == FAMG_OK; ++s) {   // interpolation/mod.rs:812-818
        famg_csr *next = nullptr;
        st = block_size == 1 ? spgemm_impl(a, p, p, omega, &next) : block_jacobi_impl(a, (int)block_size, p, &next);
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
