// dist_setup.cu -- distributed hierarchy construction: every rank builds only its row slab of
// P = (I - w D^-1 A) P0, A P and A_c = R (A P), exchanging the off-rank rows it needs.
//
// Reference: Hierarchy::coarsen (src/hierarchy.rs:190-248) -> smoothed_aggregation
// (src/interpolation/mod.rs:730-836; the three sparse expressions at :814/:938, :824-827, :828).
// The reference is shared-memory only; this is SURVEY 8(e)'s "RAP" row: the same
// products on 1-D row slabs.
//
// Per level (fine operator A_l: slab of owned rows, columns renumbered to [owned | ghost]):
//   P0   tentative prolongator of the rank's own aggregates (aggregates never straddle ranks), global
//        coarse column ids
//   P    = S(A_l, [P0 | ghost rows of P0])          ghost rows = rows of P0 owned by the neighbours that
//   AP   = A_l * [P | ghost rows of P]               A_l's ghost columns refer to (one exchange each)
//   R    = rows of P^T this rank owns: the local transpose is cut by coarse owner and the pieces are
//        exchanged; pieces concatenate in rank order => sorted rows
//   A_c  = R * [AP | ghost rows of AP]
// Products run through the single-GPU SpGEMM (spgemm.cu) on the renumbered slab; since renumbering
// never reorders the entries of a row, every output entry accumulates its contributions in ascending
// *global* inner index -- the slabs are bit-identical to the rows of the undistributed build.
//
// Exchanges are bulk-synchronous all-to-alls (xchg_*): grouped ncclSend/ncclRecv between processes,
// plain device copies when one process hosts all virtual ranks (single-GPU test mode).
#include <cmath>
#include <functional>

#include "dist_internal.cuh"

namespace famg {

constexpr int XCHG_MAX_RANKS = 16;

// ================================================================================ exchange layer
famg_status xchg_meta(famg_comm *cm, int w, const std::vector<std::vector<int64_t>> &send, std::vector<std::vector<int64_t>> &recv) {
    const int nr = cm->nranks, nl = cm->nlocal;
    recv.assign((size_t)nl, std::vector<int64_t>((size_t)nr * w, 0));
    if (nl == nr) {  // every virtual rank lives here (also the single-rank case)
        for (int s = 0; s < nr; ++s)
            for (int d = 0; d < nr; ++d)
                for (int j = 0; j < w; ++j) recv[(size_t)d][(size_t)s * w + j] = send[(size_t)s][(size_t)d * w + j];
        return FAMG_OK;
    }
    famg_ctx *ctx = cm->ctx;
    const size_t cnt = (size_t)nr * w;
    int64_t *buf = nullptr;
    const size_t bytes = sizeof(int64_t) * cnt * (nr + 1);
    FAMG_TRY(pool_alloc(ctx, bytes, (void **)&buf));
    std::vector<int64_t> all(cnt * nr);
    cudaError_t e = cudaMemcpyAsync(buf + cnt * nr, send[0].data(), sizeof(int64_t) * cnt, cudaMemcpyHostToDevice, ctx->stream);
    ncclResult_t nres = ncclSuccess;
    if (e == cudaSuccess) nres = g_nccl.AllGather(buf + cnt * nr, buf, cnt, ncclInt64, cm->comm, ctx->stream);
    if (e == cudaSuccess && nres == ncclSuccess) e = cudaMemcpyAsync(all.data(), buf, sizeof(int64_t) * cnt * nr, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    pool_free(ctx, buf, bytes);
    if (nres != ncclSuccess) FAMG_FAIL(FAMG_ERR_COMM, "setup exchange (metadata): %s", g_nccl.GetErrorString(nres));
    if (e != cudaSuccess) FAMG_FAIL(FAMG_ERR_CUDA, "setup exchange (metadata): %s", cudaGetErrorString(e));
    for (int p = 0; p < nr; ++p)
        for (int j = 0; j < w; ++j) recv[0][(size_t)p * w + j] = all[(size_t)p * cnt + (size_t)cm->rank * w + j];
    return FAMG_OK;
}

famg_status xchg_allgather_meta(famg_comm *cm, int w, const std::vector<std::vector<int64_t>> &mine, std::vector<int64_t> &all) {
    const int nr = cm->nranks, nl = cm->nlocal;
    all.assign((size_t)nr * w, 0);
    if (nl == nr) {
        for (int s = 0; s < nr; ++s)
            for (int j = 0; j < w; ++j) all[(size_t)s * w + j] = mine[(size_t)s][(size_t)j];
        return FAMG_OK;
    }
    famg_ctx *ctx = cm->ctx;
    int64_t *buf = nullptr;
    const size_t bytes = sizeof(int64_t) * (size_t)w * (nr + 1);
    FAMG_TRY(pool_alloc(ctx, bytes, (void **)&buf));
    cudaError_t e = cudaMemcpyAsync(buf + (size_t)w * nr, mine[0].data(), sizeof(int64_t) * w, cudaMemcpyHostToDevice, ctx->stream);
    ncclResult_t nres = ncclSuccess;
    if (e == cudaSuccess) nres = g_nccl.AllGather(buf + (size_t)w * nr, buf, (size_t)w, ncclInt64, cm->comm, ctx->stream);
    if (e == cudaSuccess && nres == ncclSuccess) e = cudaMemcpyAsync(all.data(), buf, sizeof(int64_t) * w * nr, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    pool_free(ctx, buf, bytes);
    if (nres != ncclSuccess) FAMG_FAIL(FAMG_ERR_COMM, "setup exchange (all-gather): %s", g_nccl.GetErrorString(nres));
    if (e != cudaSuccess) FAMG_FAIL(FAMG_ERR_CUDA, "setup exchange (all-gather): %s", cudaGetErrorString(e));
    return FAMG_OK;
}

famg_status xchg_data(famg_comm *cm, const std::vector<std::vector<XSeg>> &send, const std::vector<std::vector<XSeg>> &recv) {
    const int nr = cm->nranks, nl = cm->nlocal;
    famg_ctx *ctx = cm->ctx;
    if (nl == nr) {
        for (int s = 0; s < nr; ++s)
            for (int d = 0; d < nr; ++d) {
                const XSeg &a = send[(size_t)s][(size_t)d], &b = recv[(size_t)d][(size_t)s];
                if (a.bytes != b.bytes) FAMG_FAIL(FAMG_ERR_COMM, "internal: exchange size mismatch %d -> %d (%zu vs %zu bytes)", s, d, a.bytes, b.bytes);
                if (a.bytes) CUDA_TRY(cudaMemcpyAsync(b.p, a.p, a.bytes, cudaMemcpyDeviceToDevice, ctx->stream));
            }
        return FAMG_OK;
    }
    const int me = cm->rank;
    if (send[0][(size_t)me].bytes != recv[0][(size_t)me].bytes) FAMG_FAIL(FAMG_ERR_COMM, "internal: exchange self size mismatch");
    if (send[0][(size_t)me].bytes)
        CUDA_TRY(cudaMemcpyAsync(recv[0][(size_t)me].p, send[0][(size_t)me].p, send[0][(size_t)me].bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    NCCL_TRY(g_nccl.GroupStart());
    for (int p = 0; p < nr; ++p) {
        if (p == me) continue;
        if (send[0][(size_t)p].bytes) NCCL_TRY(g_nccl.Send(send[0][(size_t)p].p, send[0][(size_t)p].bytes, ncclInt8, p, cm->comm, ctx->stream));
        if (recv[0][(size_t)p].bytes) NCCL_TRY(g_nccl.Recv(recv[0][(size_t)p].p, recv[0][(size_t)p].bytes, ncclInt8, p, cm->comm, ctx->stream));
    }
    NCCL_TRY(g_nccl.GroupEnd());
    return FAMG_OK;
}

// ================================================================================ kernels
__global__ void oor_count_kernel(const int *__restrict__ col, int nnz, int c0, int c1, int *__restrict__ counter) {
    int local = 0;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += gridDim.x * blockDim.x) {
        const int c = col[q];
        local += (c < c0 || c >= c1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(counter, local);
}
// count, smallest and largest out-of-range column id: stats[0] += count, stats[1] = min, stats[2] = max
__global__ void oor_stats_kernel(const int *__restrict__ col, int nnz, int c0, int c1, int *__restrict__ stats) {
    int local = 0, lo = 0x7fffffff, hi = -1;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += gridDim.x * blockDim.x) {
        const int c = col[q];
        if (c < c0 || c >= c1) { ++local; lo = min(lo, c); hi = max(hi, c); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        local += __shfl_xor_sync(0xffffffffu, local, o);
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0 && local) { atomicAdd(&stats[0], local); atomicMin(&stats[1], lo); atomicMax(&stats[2], hi); }
}
__global__ void oor_mark_kernel(const int *__restrict__ col, int nnz, int c0, int c1, int gmin, int *__restrict__ flags) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += gridDim.x * blockDim.x) {
        const int c = col[q];
        if (c < c0 || c >= c1) flags[c - gmin] = 1;  // same value from every writer
    }
}
__global__ void oor_compact_kernel(const int *__restrict__ flags, const int *__restrict__ pos, int n, int gmin, int *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flags[i]) out[pos[i]] = gmin + i;
}
__global__ void oor_fill_kernel(const int *__restrict__ col, int nnz, int c0, int c1, int *__restrict__ counter, int *__restrict__ list) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += gridDim.x * blockDim.x) {
        const int c = col[q];
        if (c < c0 || c >= c1) list[atomicAdd(counter, 1)] = c;  // order is irrelevant: the host sorts and dedups
    }
}
__device__ __forceinline__ int lower_bound_dev(const int *__restrict__ a, int n, int key) {
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}
__global__ void renumber_gid_kernel(int *__restrict__ col, int nnz, int c0, int c1, const int *__restrict__ ghost, int nghost) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nnz) return;
    const int c = col[q];
    col[q] = (c >= c0 && c < c1) ? c - c0 : (c1 - c0) + lower_bound_dev(ghost, nghost, c);
}
__global__ void unrenumber_kernel(const int *__restrict__ in, int *__restrict__ out, int nnz, int c0, int nloc, const int *__restrict__ ghost) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nnz) return;
    const int c = in[q];
    out[q] = c < nloc ? c + c0 : ghost[c - nloc];
}
__global__ void add_const_kernel(int *__restrict__ v, int n, int delta) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] += delta;
}
__global__ void row_len_gather_kernel(const int *__restrict__ rp, const int *__restrict__ idx, int n, int *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) { const int i = idx[j]; out[j] = rp[i + 1] - rp[i]; }
}
// one warp per packed row
__global__ void __launch_bounds__(256) pack_rows_kernel(const int *__restrict__ rp, const int *__restrict__ col, const double *__restrict__ val,
                                                        const int *__restrict__ idx, int n, const int *__restrict__ so,
                                                        int *__restrict__ out_col, double *__restrict__ out_val) {
    const int j = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (j >= n) return;
    const int i = idx[j];
    const int b = rp[i], len = rp[i + 1] - b, o = so[j];
    for (int t = lane; t < len; t += 32) { out_col[o + t] = col[b + t]; out_val[o + t] = val[b + t]; }
}
__global__ void ext_rowptr_kernel(int *__restrict__ ext_rp, int nloc, int nnz_loc, const int *__restrict__ go, int nghost) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g <= nghost) ext_rp[nloc + g] = nnz_loc + go[g];
}
__global__ void interior_range2_kernel(const int *__restrict__ rp, const int *__restrict__ col, int nrows, int nloc_cols,
                                       int *__restrict__ lo_hi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    bool ghost = false;
    for (int q = rp[i]; q < rp[i + 1]; ++q) ghost |= col[q] >= nloc_cols;
    if (!ghost) return;
    const int mid = nrows / 2;
    if (i < mid) atomicMax(&lo_hi[0], i + 1); else atomicMin(&lo_hi[1], i);
}
__global__ void pack_vec_kernel(const double *__restrict__ x, const int *__restrict__ idx, int n, double *__restrict__ buf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) buf[i] = x[idx[i]];
}

struct MergeArgs {
    const int *seg;   // nr x (nrows + 1) row pointers as sent (absolute offsets into the sender's arrays)
    const int *rcol; const double *rval;
    int nr, nrows;
    int coff[XCHG_MAX_RANKS];  // start of peer p's entries in rcol / rval
};
__global__ void merge_len_kernel(MergeArgs m, int *__restrict__ len) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.nrows) return;
    int s = 0;
    for (int p = 0; p < m.nr; ++p) { const int *sg = m.seg + (size_t)p * (m.nrows + 1); s += sg[i + 1] - sg[i]; }
    len[i] = s;
}
__global__ void __launch_bounds__(256) merge_fill_kernel(MergeArgs m, const int *__restrict__ rp, int *__restrict__ col, double *__restrict__ val) {
    const int i = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= m.nrows) return;
    int pos = rp[i];
    for (int p = 0; p < m.nr; ++p) {  // peers in rank order: ascending global column ids
        const int *sg = m.seg + (size_t)p * (m.nrows + 1);
        const int base = sg[0], s0 = sg[i] - base, len = sg[i + 1] - sg[i];
        for (int t = lane; t < len; t += 32) { col[pos + t] = m.rcol[m.coff[p] + s0 + t]; val[pos + t] = m.rval[m.coff[p] + s0 + t]; }
        pos += len;
    }
}
__global__ void shift_rowptr_kernel(const int *__restrict__ in, int *__restrict__ out, int n, int delta) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] + delta;
}

// ================================================================================ helpers
void halo_free(HaloPlan &h) {
    cudaFree(h.d_send_idx); cudaFree(h.d_sendbuf); cudaFree(h.d_dev); cudaFree(h.d_push_map);
    h.d_send_idx = nullptr; h.d_sendbuf = nullptr; h.d_dev = nullptr; h.d_push_map = nullptr;
}
void distop_free(DistOp &o) {
    if (o.local) csr_release(o.local);
    o.local = nullptr;
    halo_free(o.halo);
    cudaFree(o.d_ghost_gid); o.d_ghost_gid = nullptr;
}

void dmat_release(famg_dmat *m) {
    if (!m) return;
    if (m->refs.fetch_sub(1) == 1) {
        for (auto &p : m->part) distop_free(p);
        delete m;
    }
}

static famg_status sync_check(famg_ctx *ctx, const char *what) {
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) FAMG_FAIL(FAMG_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return FAMG_OK;
}

static famg_dmat *dmat_new(famg_comm *cm, int64_t nrows, int64_t ncols, const std::vector<int64_t> &rsplit, const std::vector<int64_t> &csplit) {
    famg_dmat *m = new famg_dmat();
    m->comm = cm; m->nrows = nrows; m->ncols = ncols; m->rsplit = rsplit; m->csplit = csplit;
    m->part.resize((size_t)cm->nlocal);
    return m;
}

// ---------------------------------------------------------------- halo plan (collective)
famg_status dmat_finalize(famg_dmat *m, bool replicated_cols) {
    if (m->finalized) {
        if (m->replicated_cols != replicated_cols) FAMG_FAIL(FAMG_ERR_INVALID, "distributed matrix already finalized with a different column layout");
        return FAMG_OK;
    }
    famg_comm *cm = m->comm;
    famg_ctx *ctx = cm->ctx;
    const int nr = cm->nranks, nl = cm->nlocal;
    if (nr > XCHG_MAX_RANKS) FAMG_FAIL(FAMG_ERR_UNSUPPORTED, "at most %d ranks", XCHG_MAX_RANKS);
    {   // kernel plan from the statistics of the whole operator: same threads-per-row on every rank as on one GPU
        std::vector<std::vector<int64_t>> mine((size_t)nl);
        std::vector<int64_t> all;
        for (int li = 0; li < nl; ++li) mine[(size_t)li] = {m->part[(size_t)li].local->nnz};
        FAMG_TRY(xchg_allgather_meta(cm, 1, mine, all));
        int64_t total = 0;
        for (int p = 0; p < nr; ++p) total += all[(size_t)p];
        const double avg = m->nrows > 0 ? (double)total / (double)m->nrows : 0.0;
        for (int li = 0; li < nl; ++li) { m->part[(size_t)li].local->tpr = tpr_for_avg(avg); m->part[(size_t)li].local->avg_row_nnz = avg; }
    }
    for (int li = 0; li < nl; ++li) {
        DistOp &op = m->part[(size_t)li];
        HaloPlan &h = op.halo;
        h.recv_cnt.assign(nr, 0); h.recv_off.assign(nr, 0); h.send_cnt.assign(nr, 0); h.send_off.assign(nr, 0);
        op.ib = 0; op.ie = (int)op.local->nrows;
    }
    if (replicated_cols) {
        for (int li = 0; li < nl; ++li) {
            HaloPlan &h = m->part[(size_t)li].halo;
            h.nloc = (int)m->ncols; h.nghost = 0; h.any = false;
        }
        m->replicated_cols = true; m->finalized = true;
        return FAMG_OK;
    }
    // 1. ghost columns of every part: out-of-range column ids, sorted and deduplicated on the host
    //    (boundary rows only: a few planes' worth of entries)
    int *counter = nullptr;
    FAMG_TRY(pool_alloc(ctx, 256, (void **)&counter));
    auto fail = [&](famg_status s) { pool_free(ctx, counter, 256); return s; };
    for (int li = 0; li < nl; ++li) {
        DistOp &op = m->part[(size_t)li];
        HaloPlan &h = op.halo;
        const int r = cm->vrank(li);
        const int c0 = (int)m->csplit[(size_t)r], c1 = (int)m->csplit[(size_t)r + 1];
        const int nnz = (int)op.local->nnz;
        h.nloc = c1 - c0;
        // ghost columns = distinct out-of-range column ids, ascending.  They lie in a band around the owned range (a few
        // planes of a slab partition): flag them in an array over [smallest, largest], scan, compact -- on the device.
        // (A host sort of the raw list -- millions of duplicates on the wide-stencil coarse levels -- cost 0.1-0.3 s per
        // operator on 8 ranks; profiles/r2_setup_phases.md.)  Only a band wider than 2^27 falls back to the host sort.
        int stats[3] = {0, 0x7fffffff, -1};
        cudaMemcpyAsync(counter, stats, sizeof(stats), cudaMemcpyHostToDevice, ctx->stream);
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(nnz, 256), 16 * (int64_t)ctx->num_sms));
        if (nnz) {
            oor_stats_kernel<<<grid, 256, 0, ctx->stream>>>(op.local->col, nnz, c0, c1, counter);
            count_launch(ctx);
        }
        cudaMemcpyAsync(stats, counter, sizeof(stats), cudaMemcpyDeviceToHost, ctx->stream);
        famg_status st = sync_check(ctx, "halo plan (count)");
        if (st != FAMG_OK) return fail(st);
        const int cnt = stats[0];
        std::vector<int> list;
        if (cnt) {
            if (stats[1] < 0 || stats[2] >= m->ncols) return fail((set_error("column index out of range in a distributed operator"), FAMG_ERR_INVALID));
            const int64_t band = (int64_t)stats[2] - stats[1] + 1;
            if (band <= ((int64_t)1 << 27)) {
                int *flags = nullptr, *pos = nullptr, *d_list = nullptr;
                const size_t fb = sizeof(int) * (size_t)(band + 2);
                st = pool_alloc(ctx, fb, (void **)&flags);
                if (st == FAMG_OK) st = pool_alloc(ctx, fb, (void **)&pos);
                if (st != FAMG_OK) { pool_free(ctx, flags, 0); return fail(st); }
                cudaMemsetAsync(flags, 0, fb, ctx->stream);
                oor_mark_kernel<<<grid, 256, 0, ctx->stream>>>(op.local->col, nnz, c0, c1, stats[1], flags);
                count_launch(ctx);
                st = exclusive_scan_i32(ctx, flags, pos, band);
                int ng = 0;
                if (st == FAMG_OK) { cudaMemcpyAsync(&ng, pos + band, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream); st = sync_check(ctx, "halo plan (scan)"); }
                if (st == FAMG_OK) st = pool_alloc(ctx, sizeof(int) * (size_t)std::max(ng, 1), (void **)&d_list);
                if (st == FAMG_OK) {
                    oor_compact_kernel<<<(unsigned)ceil_div(band, 256), 256, 0, ctx->stream>>>(flags, pos, (int)band, stats[1], d_list);
                    count_launch(ctx);
                    list.resize((size_t)ng);
                    if (ng) cudaMemcpyAsync(list.data(), d_list, sizeof(int) * (size_t)ng, cudaMemcpyDeviceToHost, ctx->stream);
                    st = sync_check(ctx, "halo plan (ghost list)");
                }
                pool_free(ctx, flags, 0); pool_free(ctx, pos, 0); pool_free(ctx, d_list, 0);
                if (st != FAMG_OK) return fail(st);
            } else {
                list.resize((size_t)cnt);
                int *d_list = nullptr;
                st = pool_alloc(ctx, sizeof(int) * (size_t)cnt, (void **)&d_list);
                if (st != FAMG_OK) return fail(st);
                cudaMemsetAsync(counter, 0, sizeof(int), ctx->stream);
                oor_fill_kernel<<<grid, 256, 0, ctx->stream>>>(op.local->col, nnz, c0, c1, counter, d_list);
                count_launch(ctx);
                cudaMemcpyAsync(list.data(), d_list, sizeof(int) * (size_t)cnt, cudaMemcpyDeviceToHost, ctx->stream);
                st = sync_check(ctx, "halo plan (ghost list)");
                pool_free(ctx, d_list, sizeof(int) * (size_t)cnt);
                if (st != FAMG_OK) return fail(st);
                std::sort(list.begin(), list.end());
                list.erase(std::unique(list.begin(), list.end()), list.end());
            }
        }
        op.ghost_gid.swap(list);
        h.nghost = (int)op.ghost_gid.size();
        for (int p = 0; p < nr; ++p) {
            const auto lo = std::lower_bound(op.ghost_gid.begin(), op.ghost_gid.end(), (int)m->csplit[(size_t)p]);
            const auto hi = std::lower_bound(op.ghost_gid.begin(), op.ghost_gid.end(), (int)m->csplit[(size_t)p + 1]);
            h.recv_off[(size_t)p] = (int)(lo - op.ghost_gid.begin());
            h.recv_cnt[(size_t)p] = (int)(hi - lo);
        }
        if (h.recv_cnt[(size_t)r] != 0) return fail((set_error("internal: own columns flagged as ghosts"), FAMG_ERR_INVALID));
        st = dev_alloc(&op.d_ghost_gid, std::max(h.nghost, 1));
        if (st != FAMG_OK) return fail(st);
        if (h.nghost) cudaMemcpyAsync(op.d_ghost_gid, op.ghost_gid.data(), sizeof(int) * (size_t)h.nghost, cudaMemcpyHostToDevice, ctx->stream);
    }
    pool_free(ctx, counter, 256);
    // 2. every owner learns which of its entries each peer needs (request lists travel to the owners)
    std::vector<std::vector<int64_t>> req((size_t)nl), got;
    for (int li = 0; li < nl; ++li) {
        req[(size_t)li].assign((size_t)nr, 0);
        for (int p = 0; p < nr; ++p) req[(size_t)li][(size_t)p] = m->part[(size_t)li].halo.recv_cnt[(size_t)p];
    }
    FAMG_TRY(xchg_meta(cm, 1, req, got));
    std::vector<std::vector<XSeg>> snd((size_t)nl, std::vector<XSeg>((size_t)nr)), rcv((size_t)nl, std::vector<XSeg>((size_t)nr));
    for (int li = 0; li < nl; ++li) {
        DistOp &op = m->part[(size_t)li];
        HaloPlan &h = op.halo;
        h.total_send = 0;
        for (int p = 0; p < nr; ++p) { h.send_cnt[(size_t)p] = (int)got[(size_t)li][(size_t)p]; h.send_off[(size_t)p] = h.total_send; h.total_send += h.send_cnt[(size_t)p]; }
        FAMG_TRY(dev_alloc(&h.d_send_idx, std::max(h.total_send, 1)));
        FAMG_TRY(dev_alloc(&h.d_sendbuf, std::max(h.total_send, 1)));
        for (int p = 0; p < nr; ++p) {
            snd[(size_t)li][(size_t)p] = XSeg{op.d_ghost_gid + h.recv_off[(size_t)p], sizeof(int) * (size_t)h.recv_cnt[(size_t)p]};
            rcv[(size_t)li][(size_t)p] = XSeg{h.d_send_idx + h.send_off[(size_t)p], sizeof(int) * (size_t)h.send_cnt[(size_t)p]};
        }
    }
    FAMG_TRY(xchg_data(cm, snd, rcv));
    // 3. requested global ids -> local ids; renumber the slab's columns; interior row range
    int *lo_hi = nullptr;
    FAMG_TRY(pool_alloc(ctx, 256, (void **)&lo_hi));
    for (int li = 0; li < nl; ++li) {
        DistOp &op = m->part[(size_t)li];
        HaloPlan &h = op.halo;
        const int r = cm->vrank(li);
        const int c0 = (int)m->csplit[(size_t)r], c1 = (int)m->csplit[(size_t)r + 1];
        std::vector<int> sidx((size_t)h.total_send);
        if (h.total_send) {
            add_const_kernel<<<(unsigned)ceil_div(h.total_send, 256), 256, 0, ctx->stream>>>(h.d_send_idx, h.total_send, -c0);
            count_launch(ctx);
            cudaMemcpyAsync(sidx.data(), h.d_send_idx, sizeof(int) * sidx.size(), cudaMemcpyDeviceToHost, ctx->stream);
        }
        const int nnz = (int)op.local->nnz, nrows = (int)op.local->nrows;
        if (nnz) {
            renumber_gid_kernel<<<(unsigned)ceil_div(nnz, 256), 256, 0, ctx->stream>>>(op.local->col, nnz, c0, c1, op.d_ghost_gid, h.nghost);
            count_launch(ctx);
        }
        op.local->ncols = h.nloc + h.nghost;
        op.local_cols = true;
        h.any = h.total_send > 0 || h.nghost > 0;
        int init[2] = {0, nrows};
        cudaMemcpyAsync(lo_hi, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream);
        if (nrows > 0 && h.nghost > 0) {
            interior_range2_kernel<<<(unsigned)ceil_div(nrows, 256), 256, 0, ctx->stream>>>(op.local->row_ptr, op.local->col, nrows, h.nloc, lo_hi);
            count_launch(ctx);
        }
        cudaMemcpyAsync(init, lo_hi, sizeof(init), cudaMemcpyDeviceToHost, ctx->stream);
        famg_status st = sync_check(ctx, "halo plan (renumber)");
        if (st != FAMG_OK) { pool_free(ctx, lo_hi, 256); return st; }
        op.ib = init[0]; op.ie = std::max(init[0], init[1]);
        // the producer-side split: entries [0, push_lo) and [push_hi, nloc) of a vector with this column layout cover the send list
        h.push_lo = 0; h.push_hi = h.nloc;
        for (int v : sidx) {
            if (v < 0 || v >= h.nloc) { pool_free(ctx, lo_hi, 256); FAMG_FAIL(FAMG_ERR_COMM, "internal: a peer requested an entry this rank does not own"); }
            if (v < h.nloc / 2) h.push_lo = std::max(h.push_lo, v + 1); else h.push_hi = std::min(h.push_hi, v);
        }
        if (h.push_hi < h.push_lo) h.push_hi = h.push_lo;
        h.send_idx_host.swap(sidx);
    }
    pool_free(ctx, lo_hi, 256);
    m->finalized = true;
    return FAMG_OK;
}

// ---------------------------------------------------------------- ghost rows of a row-partitioned matrix
// out[li] = [ M[li] ; rows of M owned by the peers that ops[li]'s ghost columns refer to ], in ghost order.
static famg_status ext_rows(famg_comm *cm, std::vector<DistOp *> &ops, const std::vector<famg_csr *> &M, std::vector<famg_csr *> &out) {
    famg_ctx *ctx = cm->ctx;
    const int nr = cm->nranks, nl = cm->nlocal;
    out.assign((size_t)nl, nullptr);
    struct Tmp { int *slen = nullptr, *glen = nullptr, *so = nullptr, *go = nullptr, *pc = nullptr; double *pv = nullptr;
                 size_t b_s = 0, b_g = 0, b_pc = 0, b_pv = 0; std::vector<int> hso, hgo; };
    std::vector<Tmp> tmp((size_t)nl);
    auto cleanup = [&]() {
        for (auto &t : tmp) {
            pool_free(ctx, t.slen, t.b_s); pool_free(ctx, t.so, t.b_s); pool_free(ctx, t.glen, t.b_g); pool_free(ctx, t.go, t.b_g);
            pool_free(ctx, t.pc, t.b_pc); pool_free(ctx, t.pv, t.b_pv);
        }
    };
#define EX_TRY(expr) do { famg_status s__ = (expr); if (s__ != FAMG_OK) { cleanup(); for (auto &o : out) if (o) { csr_release(o); o = nullptr; } return s__; } } while (0)
    std::vector<std::vector<XSeg>> snd((size_t)nl, std::vector<XSeg>((size_t)nr)), rcv((size_t)nl, std::vector<XSeg>((size_t)nr));
    // row lengths of the requested rows travel first
    for (int li = 0; li < nl; ++li) {
        const HaloPlan &h = ops[(size_t)li]->halo;
        const famg_csr *m = M[(size_t)li];
        if (m->nrows != h.nloc) { set_error("internal: ext_rows shape mismatch (%lld rows vs %d owned columns)", (long long)m->nrows, h.nloc); EX_TRY(FAMG_ERR_INVALID); }
        Tmp &t = tmp[(size_t)li];
        t.b_s = sizeof(int) * (size_t)(h.total_send + 2); t.b_g = sizeof(int) * (size_t)(h.nghost + 2);
        EX_TRY(pool_alloc(ctx, t.b_s, (void **)&t.slen)); EX_TRY(pool_alloc(ctx, t.b_s, (void **)&t.so));
        EX_TRY(pool_alloc(ctx, t.b_g, (void **)&t.glen)); EX_TRY(pool_alloc(ctx, t.b_g, (void **)&t.go));
        if (h.total_send) {
            row_len_gather_kernel<<<(unsigned)ceil_div(h.total_send, 256), 256, 0, ctx->stream>>>(m->row_ptr, h.d_send_idx, h.total_send, t.slen);
            count_launch(ctx);
        }
        for (int p = 0; p < nr; ++p) {
            snd[(size_t)li][(size_t)p] = XSeg{t.slen + h.send_off[(size_t)p], sizeof(int) * (size_t)h.send_cnt[(size_t)p]};
            rcv[(size_t)li][(size_t)p] = XSeg{t.glen + h.recv_off[(size_t)p], sizeof(int) * (size_t)h.recv_cnt[(size_t)p]};
        }
    }
    EX_TRY(xchg_data(cm, snd, rcv));
    // offsets; pack the requested rows; allocate the extended matrices
    for (int li = 0; li < nl; ++li) {
        const HaloPlan &h = ops[(size_t)li]->halo;
        const famg_csr *m = M[(size_t)li];
        Tmp &t = tmp[(size_t)li];
        EX_TRY(exclusive_scan_i32(ctx, t.slen, t.so, h.total_send));
        EX_TRY(exclusive_scan_i32(ctx, t.glen, t.go, h.nghost));
        t.hso.resize((size_t)h.total_send + 1); t.hgo.resize((size_t)h.nghost + 1);
        cudaMemcpyAsync(t.hso.data(), t.so, sizeof(int) * t.hso.size(), cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(t.hgo.data(), t.go, sizeof(int) * t.hgo.size(), cudaMemcpyDeviceToHost, ctx->stream);
        EX_TRY(sync_check(ctx, "ghost rows (offsets)"));
        const int s_total = t.hso.back(), g_total = t.hgo.back();
        t.b_pc = sizeof(int) * (size_t)(s_total + 1); t.b_pv = sizeof(double) * (size_t)(s_total + 1);
        EX_TRY(pool_alloc(ctx, t.b_pc, (void **)&t.pc)); EX_TRY(pool_alloc(ctx, t.b_pv, (void **)&t.pv));
        if (h.total_send) {
            pack_rows_kernel<<<(unsigned)ceil_div((int64_t)h.total_send * 32, 256), 256, 0, ctx->stream>>>(m->row_ptr, m->col, m->val, h.d_send_idx,
                                                                                                          h.total_send, t.so, t.pc, t.pv);
            count_launch(ctx);
        }
        famg_csr *x = nullptr;
        EX_TRY(csr_alloc(ctx, m->nrows + h.nghost, m->ncols, m->nnz + g_total, &x));
        out[(size_t)li] = x;
        cudaMemcpyAsync(x->row_ptr, m->row_ptr, sizeof(int) * (size_t)(m->nrows + 1), cudaMemcpyDeviceToDevice, ctx->stream);
        if (m->nnz) {
            cudaMemcpyAsync(x->col, m->col, sizeof(int) * (size_t)m->nnz, cudaMemcpyDeviceToDevice, ctx->stream);
            cudaMemcpyAsync(x->val, m->val, sizeof(double) * (size_t)m->nnz, cudaMemcpyDeviceToDevice, ctx->stream);
        }
        ext_rowptr_kernel<<<(unsigned)ceil_div(h.nghost + 1, 256), 256, 0, ctx->stream>>>(x->row_ptr, (int)m->nrows, (int)m->nnz, t.go, h.nghost);
        count_launch(ctx);
    }
    for (int pass = 0; pass < 2; ++pass) {  // column ids, then values
        for (int li = 0; li < nl; ++li) {
            const HaloPlan &h = ops[(size_t)li]->halo;
            const famg_csr *m = M[(size_t)li];
            Tmp &t = tmp[(size_t)li];
            famg_csr *x = out[(size_t)li];
            const size_t es = pass == 0 ? sizeof(int) : sizeof(double);
            for (int p = 0; p < nr; ++p) {
                const int s0 = t.hso[(size_t)h.send_off[(size_t)p]], s1 = t.hso[(size_t)(h.send_off[(size_t)p] + h.send_cnt[(size_t)p])];
                const int g0 = t.hgo[(size_t)h.recv_off[(size_t)p]], g1 = t.hgo[(size_t)(h.recv_off[(size_t)p] + h.recv_cnt[(size_t)p])];
                snd[(size_t)li][(size_t)p] = XSeg{pass == 0 ? (void *)(t.pc + s0) : (void *)(t.pv + s0), es * (size_t)(s1 - s0)};
                rcv[(size_t)li][(size_t)p] = XSeg{pass == 0 ? (void *)(x->col + m->nnz + g0) : (void *)(x->val + m->nnz + g0), es * (size_t)(g1 - g0)};
            }
        }
        EX_TRY(xchg_data(cm, snd, rcv));
    }
    EX_TRY(sync_check(ctx, "ghost rows (data)"));
    cleanup();
#undef EX_TRY
    return FAMG_OK;
}

// ---------------------------------------------------------------- R = P^T, row-partitioned by coarse owner
static famg_status dist_transpose(famg_dmat *P, famg_dmat **R_out) {
    famg_comm *cm = P->comm;
    famg_ctx *ctx = cm->ctx;
    const int nr = cm->nranks, nl = cm->nlocal;
    *R_out = nullptr;
    const std::vector<int64_t> &cs = P->csplit, &fs = P->rsplit;
    std::vector<famg_csr *> PT((size_t)nl, nullptr);
    std::vector<std::vector<int>> bnd((size_t)nl, std::vector<int>((size_t)nr + 1, 0));
    famg_dmat *R = dmat_new(cm, P->ncols, P->nrows, cs, fs);
    struct Rx { int *seg = nullptr, *rcol = nullptr, *len = nullptr; double *rval = nullptr; size_t b_seg = 0, b_c = 0, b_v = 0, b_len = 0; std::vector<int> coff; int total = 0; };
    std::vector<Rx> rx((size_t)nl);
    auto cleanup = [&]() {
        for (auto &t : PT) if (t) csr_release(t);
        for (auto &x : rx) { pool_free(ctx, x.seg, x.b_seg); pool_free(ctx, x.rcol, x.b_c); pool_free(ctx, x.rval, x.b_v); pool_free(ctx, x.len, x.b_len); }
    };
#define TR_TRY(expr) do { famg_status s__ = (expr); if (s__ != FAMG_OK) { cleanup(); dmat_release(R); return s__; } } while (0)
    std::vector<std::vector<int64_t>> cnt_s((size_t)nl), cnt_r;
    for (int li = 0; li < nl; ++li) {
        const int r = cm->vrank(li);
        const famg_csr *p = P->part[(size_t)li].local;
        TR_TRY(transpose_impl(p, &PT[(size_t)li]));
        famg_csr *t = PT[(size_t)li];
        if (t->nnz) {  // local fine row ids -> global
            add_const_kernel<<<(unsigned)ceil_div(t->nnz, 256), 256, 0, ctx->stream>>>(t->col, (int)t->nnz, (int)fs[(size_t)r]);
            count_launch(ctx);
        }
        for (int q = 0; q <= nr; ++q)
            cudaMemcpyAsync(&bnd[(size_t)li][(size_t)q], t->row_ptr + cs[(size_t)q], sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        TR_TRY(sync_check(ctx, "transpose (bounds)"));
        cnt_s[(size_t)li].assign((size_t)nr, 0);
        for (int q = 0; q < nr; ++q) cnt_s[(size_t)li][(size_t)q] = bnd[(size_t)li][(size_t)q + 1] - bnd[(size_t)li][(size_t)q];
    }
    TR_TRY(xchg_meta(cm, 1, cnt_s, cnt_r));
    std::vector<std::vector<XSeg>> snd((size_t)nl, std::vector<XSeg>((size_t)nr)), rcv((size_t)nl, std::vector<XSeg>((size_t)nr));
    for (int li = 0; li < nl; ++li) {
        const int r = cm->vrank(li);
        const int nc_loc = (int)(cs[(size_t)r + 1] - cs[(size_t)r]);
        Rx &x = rx[(size_t)li];
        x.coff.assign((size_t)nr + 1, 0);
        for (int p = 0; p < nr; ++p) x.coff[(size_t)p + 1] = x.coff[(size_t)p] + (int)cnt_r[(size_t)li][(size_t)p];
        x.total = x.coff[(size_t)nr];
        x.b_seg = sizeof(int) * (size_t)nr * (size_t)(nc_loc + 1);
        x.b_c = sizeof(int) * (size_t)(x.total + 1); x.b_v = sizeof(double) * (size_t)(x.total + 1); x.b_len = sizeof(int) * (size_t)(nc_loc + 2);
        TR_TRY(pool_alloc(ctx, x.b_seg, (void **)&x.seg)); TR_TRY(pool_alloc(ctx, x.b_c, (void **)&x.rcol));
        TR_TRY(pool_alloc(ctx, x.b_v, (void **)&x.rval)); TR_TRY(pool_alloc(ctx, x.b_len, (void **)&x.len));
    }
    for (int pass = 0; pass < 3; ++pass) {  // row-pointer segments, column ids, values
        for (int li = 0; li < nl; ++li) {
            const int r = cm->vrank(li);
            const int nc_loc = (int)(cs[(size_t)r + 1] - cs[(size_t)r]);
            famg_csr *t = PT[(size_t)li];
            Rx &x = rx[(size_t)li];
            for (int p = 0; p < nr; ++p) {
                const int np = (int)(cs[(size_t)p + 1] - cs[(size_t)p]);
                const int b0 = bnd[(size_t)li][(size_t)p], b1 = bnd[(size_t)li][(size_t)p + 1];
                if (pass == 0) {
                    snd[(size_t)li][(size_t)p] = XSeg{t->row_ptr + cs[(size_t)p], sizeof(int) * (size_t)(np + 1)};
                    rcv[(size_t)li][(size_t)p] = XSeg{x.seg + (size_t)p * (size_t)(nc_loc + 1), sizeof(int) * (size_t)(nc_loc + 1)};
                } else if (pass == 1) {
                    snd[(size_t)li][(size_t)p] = XSeg{t->col + b0, sizeof(int) * (size_t)(b1 - b0)};
                    rcv[(size_t)li][(size_t)p] = XSeg{x.rcol + x.coff[(size_t)p], sizeof(int) * (size_t)cnt_r[(size_t)li][(size_t)p]};
                } else {
                    snd[(size_t)li][(size_t)p] = XSeg{t->val + b0, sizeof(double) * (size_t)(b1 - b0)};
                    rcv[(size_t)li][(size_t)p] = XSeg{x.rval + x.coff[(size_t)p], sizeof(double) * (size_t)cnt_r[(size_t)li][(size_t)p]};
                }
            }
        }
        TR_TRY(xchg_data(cm, snd, rcv));
    }
    for (int li = 0; li < nl; ++li) {
        const int r = cm->vrank(li);
        const int nc_loc = (int)(cs[(size_t)r + 1] - cs[(size_t)r]);
        Rx &x = rx[(size_t)li];
        famg_csr *rl = nullptr;
        TR_TRY(csr_alloc(ctx, nc_loc, P->nrows, x.total, &rl));
        R->part[(size_t)li].local = rl;
        MergeArgs ma{};
        ma.seg = x.seg; ma.rcol = x.rcol; ma.rval = x.rval; ma.nr = nr; ma.nrows = nc_loc;
        for (int p = 0; p < nr; ++p) ma.coff[p] = x.coff[(size_t)p];
        if (nc_loc) {
            merge_len_kernel<<<(unsigned)ceil_div(nc_loc, 256), 256, 0, ctx->stream>>>(ma, x.len);
            count_launch(ctx);
        }
        TR_TRY(exclusive_scan_i32(ctx, x.len, rl->row_ptr, nc_loc));
        if (nc_loc) {
            merge_fill_kernel<<<(unsigned)ceil_div((int64_t)nc_loc * 32, 256), 256, 0, ctx->stream>>>(ma, rl->row_ptr, rl->col, rl->val);
            count_launch(ctx);
        }
        TR_TRY(csr_finalize_plan(rl));
    }
    TR_TRY(sync_check(ctx, "transpose (merge)"));
    cleanup();
#undef TR_TRY
    *R_out = R;
    return FAMG_OK;
}

// ---------------------------------------------------------------- setup-time halo exchange of a vector
famg_status setup_halo_exchange(famg_comm *cm, std::vector<DistOp *> &ops, std::vector<double *> &x_ext) {
    famg_ctx *ctx = cm->ctx;
    const int nr = cm->nranks, nl = cm->nlocal;
    std::vector<std::vector<XSeg>> snd((size_t)nl, std::vector<XSeg>((size_t)nr)), rcv((size_t)nl, std::vector<XSeg>((size_t)nr));
    for (int li = 0; li < nl; ++li) {
        HaloPlan &h = ops[(size_t)li]->halo;
        if (h.total_send) {
            pack_vec_kernel<<<(unsigned)ceil_div(h.total_send, 256), 256, 0, ctx->stream>>>(x_ext[(size_t)li], h.d_send_idx, h.total_send, h.d_sendbuf);
            count_launch(ctx);
        }
        for (int p = 0; p < nr; ++p) {
            snd[(size_t)li][(size_t)p] = XSeg{h.d_sendbuf + h.send_off[(size_t)p], sizeof(double) * (size_t)h.send_cnt[(size_t)p]};
            rcv[(size_t)li][(size_t)p] = XSeg{x_ext[(size_t)li] + h.nloc + h.recv_off[(size_t)p], sizeof(double) * (size_t)h.recv_cnt[(size_t)p]};
        }
    }
    return xchg_data(cm, snd, rcv);
}

// gather the parts of a row-partitioned matrix into one replicated CSR per virtual rank
static famg_status dmat_gather(const famg_dmat *m, std::vector<famg_csr *> &out) {
    famg_comm *cm = m->comm;
    famg_ctx *ctx = cm->ctx;
    const int nr = cm->nranks, nl = cm->nlocal;
    out.assign((size_t)nl, nullptr);
    std::vector<std::vector<int64_t>> mine((size_t)nl);
    std::vector<int64_t> all;
    for (int li = 0; li < nl; ++li) mine[(size_t)li] = {m->part[(size_t)li].local->nnz};
    FAMG_TRY(xchg_allgather_meta(cm, 1, mine, all));
    std::vector<int64_t> noff((size_t)nr + 1, 0);
    for (int p = 0; p < nr; ++p) noff[(size_t)p + 1] = noff[(size_t)p] + all[(size_t)p];
    if (noff[(size_t)nr] >= INT32_MAX) FAMG_FAIL(FAMG_ERR_UNSUPPORTED, "gathered matrix has more than 2^31 non-zeros");
    // global column ids of every part
    std::vector<int *> gcol((size_t)nl, nullptr);
    std::vector<int *> rpbuf((size_t)nl, nullptr);
    std::vector<size_t> b_gcol((size_t)nl, 0), b_rp((size_t)nl, 0);
    auto cleanup = [&]() { for (int li = 0; li < nl; ++li) { pool_free(ctx, gcol[(size_t)li], b_gcol[(size_t)li]); pool_free(ctx, rpbuf[(size_t)li], b_rp[(size_t)li]); } };
#define GA_TRY(expr) do { famg_status s__ = (expr); if (s__ != FAMG_OK) { cleanup(); for (auto &o : out) if (o) { csr_release(o); o = nullptr; } return s__; } } while (0)
    for (int li = 0; li < nl; ++li) {
        const DistOp &op = m->part[(size_t)li];
        const int r = cm->vrank(li);
        const int nnz = (int)op.local->nnz;
        b_gcol[(size_t)li] = sizeof(int) * (size_t)(nnz + 1);
        GA_TRY(pool_alloc(ctx, b_gcol[(size_t)li], (void **)&gcol[(size_t)li]));
        if (nnz) {
            if (op.local_cols) {
                unrenumber_kernel<<<(unsigned)ceil_div(nnz, 256), 256, 0, ctx->stream>>>(op.local->col, gcol[(size_t)li], nnz, (int)m->csplit[(size_t)r], op.halo.nloc, op.d_ghost_gid);
                count_launch(ctx);
            } else {
                cudaMemcpyAsync(gcol[(size_t)li], op.local->col, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToDevice, ctx->stream);
            }
        }
        b_rp[(size_t)li] = sizeof(int) * (size_t)(m->nrows + nr + 1);
        GA_TRY(pool_alloc(ctx, b_rp[(size_t)li], (void **)&rpbuf[(size_t)li]));
        GA_TRY(csr_alloc(ctx, m->nrows, m->ncols, noff[(size_t)nr], &out[(size_t)li]));
    }
    std::vector<std::vector<XSeg>> snd((size_t)nl, std::vector<XSeg>((size_t)nr)), rcv((size_t)nl, std::vector<XSeg>((size_t)nr));
    for (int pass = 0; pass < 3; ++pass) {
        for (int li = 0; li < nl; ++li) {
            const DistOp &op = m->part[(size_t)li];
            famg_csr *g = out[(size_t)li];
            for (int p = 0; p < nr; ++p) {
                const int np = (int)(m->rsplit[(size_t)p + 1] - m->rsplit[(size_t)p]);
                if (pass == 0) {
                    snd[(size_t)li][(size_t)p] = XSeg{op.local->row_ptr, sizeof(int) * (size_t)(op.local->nrows + 1)};
                    rcv[(size_t)li][(size_t)p] = XSeg{rpbuf[(size_t)li] + m->rsplit[(size_t)p] + p, sizeof(int) * (size_t)(np + 1)};
                } else if (pass == 1) {
                    snd[(size_t)li][(size_t)p] = XSeg{gcol[(size_t)li], sizeof(int) * (size_t)op.local->nnz};
                    rcv[(size_t)li][(size_t)p] = XSeg{g->col + noff[(size_t)p], sizeof(int) * (size_t)all[(size_t)p]};
                } else {
                    snd[(size_t)li][(size_t)p] = XSeg{op.local->val, sizeof(double) * (size_t)op.local->nnz};
                    rcv[(size_t)li][(size_t)p] = XSeg{g->val + noff[(size_t)p], sizeof(double) * (size_t)all[(size_t)p]};
                }
            }
        }
        GA_TRY(xchg_data(cm, snd, rcv));
    }
    for (int li = 0; li < nl; ++li) {
        famg_csr *g = out[(size_t)li];
        for (int p = 0; p < nr; ++p) {
            const int np = (int)(m->rsplit[(size_t)p + 1] - m->rsplit[(size_t)p]);
            // rows of peer p: its local row pointers shifted by the non-zeros of the peers before it
            shift_rowptr_kernel<<<(unsigned)ceil_div(np + 1, 256), 256, 0, ctx->stream>>>(rpbuf[(size_t)li] + m->rsplit[(size_t)p] + p,
                                                                                           g->row_ptr + m->rsplit[(size_t)p], np + 1, (int)noff[(size_t)p]);
            count_launch(ctx);
        }
        GA_TRY(csr_finalize_plan(g));
    }
    GA_TRY(sync_check(ctx, "gather"));
    cleanup();
#undef GA_TRY
    return FAMG_OK;
}

}  // namespace famg

using namespace famg;

extern "C" {

famg_status famg_comm_create_sim(famg_ctx *ctx, int nranks, famg_comm **out) {
    if (!ctx || !out || nranks < 1 || nranks > XCHG_MAX_RANKS) FAMG_FAIL(FAMG_ERR_INVALID, "bad argument (1 <= nranks <= %d)", XCHG_MAX_RANKS);
    *out = nullptr;
    CUDA_TRY(cudaSetDevice(ctx->device));
    famg_comm *c = new famg_comm();
    c->ctx = ctx; c->nranks = nranks; c->rank = 0; c->nlocal = nranks;
    cudaEventCreateWithFlags(&c->ev_packed, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_packed_e, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_halo_e, cudaEventDisableTiming);
    *out = c;
    return FAMG_OK;
}

famg_status famg_comm_dims(const famg_comm *c, int *nranks, int *rank, int *nlocal) {
    if (!c) FAMG_FAIL(FAMG_ERR_INVALID, "null comm");
    if (nranks) *nranks = c->nranks;
    if (rank) *rank = c->rank;
    if (nlocal) *nlocal = c->nlocal;
    return FAMG_OK;
}

famg_status famg_comm_allgatherv_f64(famg_comm *c, const double *const *local, const int64_t *counts, double *const *out) {
    if (!c || !local || !counts || !out) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    famg_ctx *ctx = c->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int nr = c->nranks, nl = c->nlocal;
    std::vector<std::vector<int64_t>> mine((size_t)nl);
    std::vector<int64_t> all;
    for (int li = 0; li < nl; ++li) mine[(size_t)li] = {counts[li]};
    FAMG_TRY(xchg_allgather_meta(c, 1, mine, all));
    std::vector<int64_t> off((size_t)nr + 1, 0);
    for (int p = 0; p < nr; ++p) off[(size_t)p + 1] = off[(size_t)p] + all[(size_t)p];
    const int64_t total = off[(size_t)nr];
    std::vector<double *> dl((size_t)nl, nullptr), dg((size_t)nl, nullptr);
    std::vector<size_t> bl((size_t)nl, 0);
    const size_t bg = sizeof(double) * (size_t)(total + 1);
    famg_status st = FAMG_OK;
    for (int li = 0; li < nl && st == FAMG_OK; ++li) {
        bl[(size_t)li] = sizeof(double) * (size_t)(counts[li] + 1);
        st = pool_alloc(ctx, bl[(size_t)li], (void **)&dl[(size_t)li]);
        if (st == FAMG_OK) st = pool_alloc(ctx, bg, (void **)&dg[(size_t)li]);
        if (st == FAMG_OK && counts[li]) cudaMemcpyAsync(dl[(size_t)li], local[li], sizeof(double) * (size_t)counts[li], cudaMemcpyHostToDevice, ctx->stream);
    }
    if (st == FAMG_OK) {
        std::vector<std::vector<XSeg>> snd((size_t)nl, std::vector<XSeg>((size_t)nr)), rcv((size_t)nl, std::vector<XSeg>((size_t)nr));
        for (int li = 0; li < nl; ++li)
            for (int p = 0; p < nr; ++p) {
                snd[(size_t)li][(size_t)p] = XSeg{dl[(size_t)li], sizeof(double) * (size_t)counts[li]};
                rcv[(size_t)li][(size_t)p] = XSeg{dg[(size_t)li] + off[(size_t)p], sizeof(double) * (size_t)all[(size_t)p]};
            }
        st = xchg_data(c, snd, rcv);
    }
    for (int li = 0; li < nl && st == FAMG_OK; ++li)
        if (total) cudaMemcpyAsync(out[li], dg[(size_t)li], sizeof(double) * (size_t)total, cudaMemcpyDeviceToHost, ctx->stream);
    if (st == FAMG_OK) st = sync_check(ctx, "allgatherv");
    for (int li = 0; li < nl; ++li) { pool_free(ctx, dl[(size_t)li], bl[(size_t)li]); pool_free(ctx, dg[(size_t)li], bg); }
    return st;
}

famg_status famg_dmat_create(famg_comm *c, famg_csr *const *slabs, int64_t ncols_global, const int64_t *col_split, famg_dmat **out) {
    if (!c || !slabs || !out) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    *out = nullptr;
    famg_ctx *ctx = c->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int nr = c->nranks, nl = c->nlocal;
    if (nr > XCHG_MAX_RANKS) FAMG_FAIL(FAMG_ERR_UNSUPPORTED, "at most %d ranks", XCHG_MAX_RANKS);
    std::vector<std::vector<int64_t>> mine((size_t)nl);
    for (int li = 0; li < nl; ++li) {
        if (!slabs[li] || slabs[li]->ctx != ctx) FAMG_FAIL(FAMG_ERR_INVALID, "slab %d is null or lives on another context", li);
        if (slabs[li]->ncols != ncols_global) FAMG_FAIL(FAMG_ERR_INVALID, "slab %d has %lld columns, expected the global %lld", li, (long long)slabs[li]->ncols, (long long)ncols_global);
        mine[(size_t)li] = {slabs[li]->nrows};
    }
    std::vector<int64_t> all;
    FAMG_TRY(xchg_allgather_meta(c, 1, mine, all));
    std::vector<int64_t> rs((size_t)nr + 1, 0), cs;
    for (int p = 0; p < nr; ++p) rs[(size_t)p + 1] = rs[(size_t)p] + all[(size_t)p];
    if (col_split) {
        cs.assign(col_split, col_split + nr + 1);
        bool ok = cs[0] == 0 && cs[(size_t)nr] == ncols_global;
        for (int p = 0; p < nr; ++p) ok = ok && cs[(size_t)p] <= cs[(size_t)p + 1];
        if (!ok) FAMG_FAIL(FAMG_ERR_INVALID, "col_split does not partition the columns");
    } else {
        if (rs[(size_t)nr] != ncols_global) FAMG_FAIL(FAMG_ERR_INVALID, "square operator expected (rows %lld, columns %lld): pass col_split", (long long)rs[(size_t)nr], (long long)ncols_global);
        cs = rs;
    }
    if (rs[(size_t)nr] >= INT32_MAX || ncols_global >= INT32_MAX) FAMG_FAIL(FAMG_ERR_UNSUPPORTED, "global dimensions exceed 32-bit indices");
    famg_dmat *m = dmat_new(c, rs[(size_t)nr], ncols_global, rs, cs);
    for (int li = 0; li < nl; ++li) { slabs[li]->refs.fetch_add(1); m->part[(size_t)li].local = slabs[li]; }
    *out = m;
    return FAMG_OK;
}

famg_status famg_dmat_finalize(famg_dmat *m, int replicated_cols) {
    if (!m) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(m->comm->ctx->device));
    return dmat_finalize(m, replicated_cols != 0);
}

famg_status famg_dmat_retain(famg_dmat *m) {
    if (!m) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    m->refs.fetch_add(1);
    return FAMG_OK;
}

famg_status famg_dmat_destroy(famg_dmat *m) {
    if (!m) return FAMG_OK;
    cudaSetDevice(m->comm->ctx->device);
    dmat_release(m);
    return FAMG_OK;
}

famg_status famg_dmat_info(const famg_dmat *m, int64_t *nrows, int64_t *ncols, int64_t *row_split, int64_t *col_split) {
    if (!m) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    if (nrows) *nrows = m->nrows;
    if (ncols) *ncols = m->ncols;
    if (row_split) std::copy(m->rsplit.begin(), m->rsplit.end(), row_split);
    if (col_split) std::copy(m->csplit.begin(), m->csplit.end(), col_split);
    return FAMG_OK;
}

famg_status famg_dmat_local(const famg_dmat *m, int local_index, int global_cols, famg_csr **out) {
    if (!m || !out || local_index < 0 || local_index >= m->comm->nlocal) FAMG_FAIL(FAMG_ERR_INVALID, "bad argument");
    *out = nullptr;
    famg_ctx *ctx = m->comm->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const DistOp &op = m->part[(size_t)local_index];
    if (!global_cols || !op.local_cols) {
        if (global_cols && m->finalized && !m->replicated_cols && !op.local_cols) FAMG_FAIL(FAMG_ERR_INVALID, "internal: finalized without renumbering");
        op.local->refs.fetch_add(1);
        *out = op.local;
        return FAMG_OK;
    }
    const int r = m->comm->vrank(local_index);
    famg_csr *g = nullptr;
    FAMG_TRY(csr_alloc(ctx, op.local->nrows, m->ncols, op.local->nnz, &g));
    const int nnz = (int)op.local->nnz;
    cudaMemcpyAsync(g->row_ptr, op.local->row_ptr, sizeof(int) * (size_t)(op.local->nrows + 1), cudaMemcpyDeviceToDevice, ctx->stream);
    if (nnz) {
        unrenumber_kernel<<<(unsigned)ceil_div(nnz, 256), 256, 0, ctx->stream>>>(op.local->col, g->col, nnz, (int)m->csplit[(size_t)r], op.halo.nloc, op.d_ghost_gid);
        count_launch(ctx);
        cudaMemcpyAsync(g->val, op.local->val, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToDevice, ctx->stream);
    }
    famg_status st = csr_finalize_plan(g);
    if (st != FAMG_OK) { csr_release(g); return st; }
    *out = g;
    return FAMG_OK;
}

famg_status famg_dmat_gather(const famg_dmat *m, famg_csr **out) {
    if (!m || !out) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(m->comm->ctx->device));
    std::vector<famg_csr *> g;
    FAMG_TRY(dmat_gather(m, g));
    for (int li = 0; li < m->comm->nlocal; ++li) out[li] = g[(size_t)li];
    return FAMG_OK;
}

// One coarsening step on row slabs (smoothed_aggregation, interpolation/mod.rs:730-836, block_size 1, one
// near-null vector).  `a` must be finalized (halo plan built).  Aggregates are local to the rank: agg_nodes are
// local row ids; coarse ids are assigned rank after rank.  make_p0(li, &p0): the rank's tentative prolongator with
// local coarse column ids (host or device construction).
static famg_status dist_coarsen_impl(famg_dmat *a, const std::vector<int64_t> &n_aggs, const std::function<famg_status(int, famg_csr **)> &make_p0,
                                     int smoothing_steps, double omega, famg_dmat **p_out, famg_dmat **r_out, famg_dmat **ac_out) {
    *p_out = *r_out = *ac_out = nullptr;
    famg_comm *cm = a->comm;
    famg_ctx *ctx = cm->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (!a->finalized || a->replicated_cols) FAMG_FAIL(FAMG_ERR_INVALID, "the fine operator must be finalized (famg_dmat_finalize) before coarsening");
    if (a->nrows != a->ncols || a->rsplit != a->csplit) FAMG_FAIL(FAMG_ERR_INVALID, "the fine operator must be square with conformal row / column splits");
    const int nr = cm->nranks, nl = cm->nlocal;
    // coarse numbering: rank after rank
    std::vector<std::vector<int64_t>> mine((size_t)nl);
    std::vector<int64_t> all;
    for (int li = 0; li < nl; ++li) mine[(size_t)li] = {n_aggs[(size_t)li]};
    FAMG_TRY(xchg_allgather_meta(cm, 1, mine, all));
    std::vector<int64_t> cs((size_t)nr + 1, 0);
    for (int p = 0; p < nr; ++p) cs[(size_t)p + 1] = cs[(size_t)p] + all[(size_t)p];
    const int64_t nc = cs[(size_t)nr];
    if (nc >= INT32_MAX) FAMG_FAIL(FAMG_ERR_UNSUPPORTED, "coarse dimension exceeds 32-bit indices");

    std::vector<DistOp *> aops((size_t)nl);
    for (int li = 0; li < nl; ++li) aops[(size_t)li] = &a->part[(size_t)li];
    famg_dmat *P = dmat_new(cm, a->nrows, nc, a->rsplit, cs), *R = nullptr, *AC = nullptr;
    std::vector<famg_csr *> cur((size_t)nl, nullptr), ext, ap((size_t)nl, nullptr);
    auto drop = [&](std::vector<famg_csr *> &v) { for (auto &m : v) if (m) { csr_release(m); m = nullptr; } };
    auto fail = [&](famg_status s) {
        drop(cur); drop(ext); drop(ap);
        dmat_release(P); if (R) dmat_release(R); if (AC) dmat_release(AC);
        return s;
    };
#define CO_TRY(expr) do { famg_status s__ = (expr); if (s__ != FAMG_OK) return fail(s__); } while (0)
    // tentative prolongator of the rank's own aggregates (interpolation/mod.rs:747-809), global coarse ids
    for (int li = 0; li < nl; ++li) {
        PhaseTimer pt(ctx, "tentative P");
        const int r = cm->vrank(li);
        const int64_t nloc = a->rsplit[(size_t)r + 1] - a->rsplit[(size_t)r];
        famg_csr *p0 = nullptr;
        CO_TRY(make_p0(li, &p0));
        cur[(size_t)li] = p0;
        if (p0->nrows != nloc || p0->ncols != n_aggs[(size_t)li]) { set_error("tentative prolongator of rank %d has the wrong shape", r); CO_TRY(FAMG_ERR_INVALID); }
        if (p0->nnz && cs[(size_t)r]) {
            add_const_kernel<<<(unsigned)ceil_div(p0->nnz, 256), 256, 0, ctx->stream>>>(p0->col, (int)p0->nnz, (int)cs[(size_t)r]);
            count_launch(ctx);
        }
        p0->ncols = nc;
    }
    // P <- (I - w D^-1 A) P  (interpolation/mod.rs:812-818, 927-946)
    for (int s = 0; s < smoothing_steps; ++s) {
        { PhaseTimer pt(ctx, "ghost rows of P0"); CO_TRY(ext_rows(cm, aops, cur, ext)); }
        PhaseTimer pt(ctx, "SpGEMM (I - w D^-1 A) P0");
        for (int li = 0; li < nl; ++li) {
            famg_csr *next = nullptr;
            CO_TRY(spgemm_impl(aops[(size_t)li]->local, ext[(size_t)li], cur[(size_t)li], omega, &next));
            csr_release(cur[(size_t)li]);
            cur[(size_t)li] = next;
        }
        drop(ext);
    }
    for (int li = 0; li < nl; ++li) { P->part[(size_t)li].local = cur[(size_t)li]; cur[(size_t)li] = nullptr; }
    // A P  (:828 inner)
    {
        std::vector<famg_csr *> pl((size_t)nl);
        for (int li = 0; li < nl; ++li) pl[(size_t)li] = P->part[(size_t)li].local;
        { PhaseTimer pt(ctx, "ghost rows of P"); CO_TRY(ext_rows(cm, aops, pl, ext)); }
        PhaseTimer pt(ctx, "SpGEMM A P");
        for (int li = 0; li < nl; ++li) CO_TRY(spgemm_impl(aops[(size_t)li]->local, ext[(size_t)li], nullptr, 0.0, &ap[(size_t)li]));
        drop(ext);
    }
    // R = P^T by coarse owner (:824-827), then its halo plan over the fine split
    { PhaseTimer pt(ctx, "R = P^T (transpose + exchange)"); CO_TRY(dist_transpose(P, &R)); }
    { PhaseTimer pt(ctx, "halo plan of R"); CO_TRY(dmat_finalize(R, false)); }
    // A_c = R (A P)  (:828 outer)
    {
        std::vector<DistOp *> rops((size_t)nl);
        for (int li = 0; li < nl; ++li) rops[(size_t)li] = &R->part[(size_t)li];
        { PhaseTimer pt(ctx, "ghost rows of AP"); CO_TRY(ext_rows(cm, rops, ap, ext)); }
        AC = dmat_new(cm, nc, nc, cs, cs);
        PhaseTimer pt(ctx, "SpGEMM R (AP)");
        for (int li = 0; li < nl; ++li) CO_TRY(spgemm_impl(rops[(size_t)li]->local, ext[(size_t)li], nullptr, 0.0, &AC->part[(size_t)li].local));
        drop(ext);
    }
    drop(ap);
#undef CO_TRY
    *p_out = P; *r_out = R; *ac_out = AC;
    return FAMG_OK;
}

famg_status famg_dist_coarsen(famg_dmat *a, const int64_t *n_aggs, const uint64_t *const *agg_ptr, const uint64_t *const *agg_nodes,
                              const double *const *near_null, int smoothing_steps, double omega, famg_dmat **p_out, famg_dmat **r_out,
                              famg_dmat **ac_out, double *const *coarse_nn) {
    if (!a || !n_aggs || !agg_ptr || !agg_nodes || !near_null || !p_out || !r_out || !ac_out || !coarse_nn || smoothing_steps < 0)
        FAMG_FAIL(FAMG_ERR_INVALID, "bad argument");
    famg_comm *cm = a->comm;
    std::vector<int64_t> na(n_aggs, n_aggs + cm->nlocal);
    auto make = [&](int li, famg_csr **p0) -> famg_status {
        const int r = cm->vrank(li);
        const int64_t nloc = a->rsplit[(size_t)r + 1] - a->rsplit[(size_t)r];
        return famg_tentative_p(cm->ctx, nloc, 1, 1, 1, near_null[li], std::max<int64_t>(nloc, 1), n_aggs[li], agg_ptr[li], agg_nodes[li], p0, coarse_nn[li]);
    };
    return dist_coarsen_impl(a, na, make, smoothing_steps, omega, p_out, r_out, ac_out);
}

famg_status famg_dist_coarsen_dev(famg_dmat *a, famg_partition *const *parts, const famg_vec *const *near_null, int smoothing_steps,
                                  double omega, famg_dmat **p_out, famg_dmat **r_out, famg_dmat **ac_out, famg_vec *const *coarse_nn) {
    if (!a || !parts || !near_null || !p_out || !r_out || !ac_out || !coarse_nn || smoothing_steps < 0) FAMG_FAIL(FAMG_ERR_INVALID, "bad argument");
    famg_comm *cm = a->comm;
    std::vector<int64_t> na((size_t)cm->nlocal);
    for (int li = 0; li < cm->nlocal; ++li) {
        if (!parts[li] || !near_null[li] || !coarse_nn[li]) FAMG_FAIL(FAMG_ERR_INVALID, "null per-rank argument");
        na[(size_t)li] = parts[li]->n_aggs;
    }
    auto make = [&](int li, famg_csr **p0) -> famg_status { return famg_tentative_p_dev(parts[li], near_null[li], p0, coarse_nn[li]); };
    return dist_coarsen_impl(a, na, make, smoothing_steps, omega, p_out, r_out, ac_out);
}

// Coarse near-null of a distributed level (hierarchy.rs:217-228): three steps of the L1 stationary iteration
// (smoothers.rs:146-159, literal) and the thin QR of the single column, i.e. division by its 2-norm.  The sum
// of squares is accumulated sequentially rank after rank -- the same order as the undistributed host loop
// (famg_thin_q), so the result is bit-identical to the single-GPU build.
famg_status famg_dist_smooth_near_null(famg_dmat *a, int iters, double *const *nn) {
    if (!a || !nn || iters < 1) FAMG_FAIL(FAMG_ERR_INVALID, "bad argument");
    famg_comm *cm = a->comm;
    famg_ctx *ctx = cm->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (!a->finalized || a->replicated_cols || a->nrows != a->ncols) FAMG_FAIL(FAMG_ERR_INVALID, "finalized square distributed operator required");
    const int nr = cm->nranks, nl = cm->nlocal;
    std::vector<DistOp *> ops((size_t)nl);
    std::vector<double *> cur((size_t)nl, nullptr), oth((size_t)nl, nullptr), dd((size_t)nl, nullptr);
    std::vector<size_t> bytes((size_t)nl, 0);
    famg_status st = FAMG_OK;
    for (int li = 0; li < nl && st == FAMG_OK; ++li) {
        ops[(size_t)li] = &a->part[(size_t)li];
        const HaloPlan &h = ops[(size_t)li]->halo;
        bytes[(size_t)li] = sizeof(double) * (size_t)(h.nloc + h.nghost + 2);
        st = pool_alloc(ctx, bytes[(size_t)li], (void **)&cur[(size_t)li]);
        if (st == FAMG_OK) st = pool_alloc(ctx, bytes[(size_t)li], (void **)&oth[(size_t)li]);
        if (st == FAMG_OK) st = pool_alloc(ctx, bytes[(size_t)li], (void **)&dd[(size_t)li]);
        if (st != FAMG_OK) break;
        cudaMemsetAsync(cur[(size_t)li], 0, bytes[(size_t)li], ctx->stream);
        cudaMemsetAsync(oth[(size_t)li], 0, bytes[(size_t)li], ctx->stream);
        if (h.nloc) cudaMemcpyAsync(oth[(size_t)li], nn[li], sizeof(double) * (size_t)h.nloc, cudaMemcpyHostToDevice, ctx->stream);
        st = diag_from_rows(ops[(size_t)li]->local, FAMG_DIAG_L1, 0.0, dd[(size_t)li]);
        // x = M^-1 rhs
        if (st == FAMG_OK) st = vec_scale_rows(ctx, dd[(size_t)li], oth[(size_t)li], 0, cur[(size_t)li], 0, h.nloc, 1);
    }
    for (int it = 1; it < iters && st == FAMG_OK; ++it) {
        st = setup_halo_exchange(cm, ops, cur);
        for (int li = 0; li < nl && st == FAMG_OK; ++li) {
            SpmvArgs g; g.a = ops[(size_t)li]->local; g.epi = EPI_SI; g.x = cur[(size_t)li]; g.y = oth[(size_t)li]; g.d = dd[(size_t)li]; g.k = 1;
            st = spmv_launch(g);
        }
        cur.swap(oth);
    }
    for (int li = 0; li < nl && st == FAMG_OK; ++li) {
        const int nloc = ops[(size_t)li]->halo.nloc;
        if (nloc) cudaMemcpyAsync(nn[li], cur[(size_t)li], sizeof(double) * (size_t)nloc, cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (st == FAMG_OK) st = sync_check(ctx, "near-null smoothing");
    for (int li = 0; li < nl; ++li) { pool_free(ctx, cur[(size_t)li], bytes[(size_t)li]); pool_free(ctx, oth[(size_t)li], bytes[(size_t)li]); pool_free(ctx, dd[(size_t)li], bytes[(size_t)li]); }
    FAMG_TRY(st);
    // thin Q of one column: v / ||v||, the sum of squares chained through the ranks in order
    double running = 0.0;
    for (int p = 0; p < nr; ++p) {
        std::vector<std::vector<int64_t>> mine((size_t)nl, std::vector<int64_t>(1, 0));
        for (int li = 0; li < nl; ++li) {
            if (cm->vrank(li) != p) continue;
            const int nloc = ops[(size_t)li]->halo.nloc;
            double s = running;
            for (int i = 0; i < nloc; ++i) s += nn[li][i] * nn[li][i];
            memcpy(&mine[(size_t)li][0], &s, sizeof(double));
        }
        std::vector<int64_t> all;
        FAMG_TRY(xchg_allgather_meta(cm, 1, mine, all));
        memcpy(&running, &all[(size_t)p], sizeof(double));
    }
    const double norm = sqrt(running);
    if (!(norm > 0)) FAMG_FAIL(FAMG_ERR_NUMERIC, "thin_q: column 0 is linearly dependent");
    for (int li = 0; li < nl; ++li) {
        const int nloc = ops[(size_t)li]->halo.nloc;
        for (int i = 0; i < nloc; ++i) nn[li][i] /= norm;
    }
    return FAMG_OK;
}

// device near-null columns: staged through the host routine (the chained sum of squares is a host loop by construction;
// the columns of a coarse level are a few MB)
famg_status famg_dist_smooth_near_null_dev(famg_dmat *a, int iters, famg_vec *const *nn) {
    if (!a || !nn) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    const int nl = a->comm->nlocal;
    std::vector<std::vector<double>> host((size_t)nl);
    std::vector<double *> ptrs((size_t)nl);
    for (int li = 0; li < nl; ++li) {
        if (!nn[li] || nn[li]->ncols != 1 || nn[li]->nrows != a->part[(size_t)li].halo.nloc) FAMG_FAIL(FAMG_ERR_INVALID, "near-null column %d has the wrong shape", li);
        host[(size_t)li].resize((size_t)std::max<int64_t>(nn[li]->nrows, 1));
        ptrs[(size_t)li] = host[(size_t)li].data();
        if (nn[li]->nrows) FAMG_TRY(famg_vec_download(nn[li], ptrs[(size_t)li], nn[li]->nrows));
    }
    FAMG_TRY(famg_dist_smooth_near_null(a, iters, ptrs.data()));
    for (int li = 0; li < nl; ++li)
        if (nn[li]->nrows) FAMG_TRY(famg_vec_upload(nn[li], ptrs[(size_t)li], nn[li]->nrows));
    return FAMG_OK;
}

}  // extern "C"
