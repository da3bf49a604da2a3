// multigrid.cu -- device-resident Multigrid: level containers and the mu-cycle.
// Reference: src/preconditioners/multigrid.rs:171-249 (type, new, add_level, with_*),
// :251-267 (init_cycle), :269-380 (cycle), :407-424 (smooth), :469-473 (apply).
//
// Per level and visit the reference runs  smooth / A v / f - work / R work / recurse x mu /
// P v_c / v += work / smooth, allocating and zeroing fresh matrices every time.  Here each level
// owns a workspace allocated once, and a visit is at most 4 fused kernels + 1 streaming kernel:
//   v  = d .* f                      first pre-sweep from a zero guess (A*0 is never formed)
//   v' = v + d .* (f - A v)          remaining sweeps, one read of A each (EPI_SMOOTH)
//   w  = f - A v                     (EPI_RESID)
//   f_c = R w                        (EPI_SPMV)
//   v += P v_c                       (EPI_ADD)
// The whole cycle is captured once into a CUDA graph per (out, rhs, k) and replayed, which removes
// the launch latency that otherwise dominates the coarse levels.
#include "common.cuh"

#include "mg_internal.cuh"

namespace famg {

static void mg_free_workspace(famg_mg *mg) {
    for (auto &g : mg->graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
    mg->graphs.clear();
    for (auto &l : mg->lv) {
        cudaFree(l.x); cudaFree(l.b); cudaFree(l.t); cudaFree(l.z);
        l.x = l.b = l.t = l.z = nullptr;
    }
    mg->ws_k = 0;
}

famg_status mg_ensure_workspace(famg_mg *mg, int k) {
    if (k <= mg->ws_k) return FAMG_OK;
    cudaStreamSynchronize(mg->ctx->stream);
    mg_free_workspace(mg);
    for (size_t i = 0; i < mg->lv.size(); ++i) {
        MgLevel &l = mg->lv[i];
        const int64_t n = l.a->nrows;
        l.ld = (n + 1) & ~(int64_t)1;
        const int64_t cnt = l.ld * k + 2;
        FAMG_TRY(dev_alloc(&l.t, cnt));
        if (i > 0) { FAMG_TRY(dev_alloc(&l.x, cnt)); FAMG_TRY(dev_alloc(&l.b, cnt)); }
        if (l.s->kind != SM_DIAG) FAMG_TRY(dev_alloc(&l.z, cnt));
    }
    mg->ws_k = k;
    return FAMG_OK;
}

// One visit of `level`.  The result is written to (va, lda); `zero_guess` says the incoming
// iterate is identically zero (first visit), otherwise (va, lda) holds it.
famg_status mg_cycle(famg_mg *mg, size_t level, double *va, int64_t lda, const double *f, int64_t ldf, int k,
                            bool zero_guess, double *dot_partials, int *dot_count) {
    famg_ctx *ctx = mg->ctx;
    MgLevel &L = mg->lv[level];
    const int64_t n = L.a->nrows;
    if (dot_count) *dot_count = 0;
    if (level + 1 == mg->lv.size()) {
        // smoother.apply(v, f): v = M^-1 f whatever v held (multigrid.rs:292)
        return smoother_apply_dev(L.s, f, ldf, va, lda, k);
    }
    const bool diag = L.s->kind == SM_DIAG;
    double *cur = va, *oth = L.t;
    int64_t ldc = lda, ldo = L.ld;
    auto sweep = [&](bool with_dot = false) -> famg_status {
        if (diag) {
            SpmvArgs g; g.a = L.a; g.epi = EPI_SMOOTH; g.x = cur; g.ldx = ldc; g.y = oth; g.ldy = ldo; g.b = f; g.ldb = ldf;
            g.d = L.s->d; g.k = k;
            if (with_dot) g.dot_partials = dot_partials;
            FAMG_TRY(spmv_launch(g, with_dot ? dot_count : nullptr));
            std::swap(cur, oth); std::swap(ldc, ldo);
            return FAMG_OK;
        }
        SpmvArgs g; g.a = L.a; g.epi = EPI_RESID; g.x = cur; g.ldx = ldc; g.y = oth; g.ldy = ldo; g.b = f; g.ldb = ldf; g.k = k;
        FAMG_TRY(spmv_launch(g));
        FAMG_TRY(smoother_apply_dev(L.s, oth, ldo, L.z, L.ld, k));
        for (int c = 0; c < k; ++c) FAMG_TRY(vec_add_inplace(ctx, cur + c * ldc, L.z + c * L.ld, n));
        return FAMG_OK;
    };
    int pre = mg->nu;
    if (zero_guess) {
        if (diag) {
            // pick the start buffer so that after all ping-pong sweeps the iterate sits in va
            const int swaps = (mg->nu - 1) + mg->nu;
            if (swaps & 1) { std::swap(cur, oth); std::swap(ldc, ldo); }
        }
        // first sweep from zero: v = 0 + M^-1 (f - A 0) = M^-1 f
        FAMG_TRY(smoother_apply_dev(L.s, f, ldf, cur, ldc, k));
        pre -= 1;
    }
    for (int i = 0; i < pre; ++i) FAMG_TRY(sweep());                                 // multigrid.rs:314-322
    MgLevel &C = mg->lv[level + 1];
    {
        SpmvArgs g; g.a = L.a; g.epi = EPI_RESID; g.x = cur; g.ldx = ldc; g.y = oth; g.ldy = ldo; g.b = f; g.ldb = ldf; g.k = k;
        FAMG_TRY(spmv_launch(g));                                                    // :341-342
        SpmvArgs r; r.a = C.r; r.epi = EPI_SPMV; r.x = oth; r.ldx = ldo; r.y = C.b; r.ldy = C.ld; r.k = k;
        FAMG_TRY(spmv_launch(r));                                                    // :343
    }
    for (int m = 0; m < mg->mu; ++m)                                                 // :345-347
        FAMG_TRY(mg_cycle(mg, level + 1, C.x, C.ld, C.b, C.ld, k, m == 0));
    {
        SpmvArgs g; g.a = C.p; g.epi = EPI_ADD; g.x = C.x; g.ldx = C.ld; g.y = cur; g.ldy = ldc; g.k = k;
        FAMG_TRY(spmv_launch(g));                                                    // :349-350
    }
    for (int i = 0; i < mg->nu; ++i)                                                 // :361-369
        FAMG_TRY(sweep(i + 1 == mg->nu && diag && k == 1 && dot_partials != nullptr && dot_count != nullptr));
    if (cur != va) FAMG_FAIL(FAMG_ERR_INVALID, "internal: multigrid ping-pong parity broken");
    return FAMG_OK;
}

famg_status mg_apply_dev_raw(famg_mg *mg, double *out, int64_t ldo, const double *rhs, int64_t ldr, int k,
                             double *dot_partials, int *dot_count) {
    famg_ctx *ctx = mg->ctx;
    FAMG_TRY(mg_ensure_workspace(mg, k));
    if (dot_count) *dot_count = 0;
    if (mg->lv.size() < 2 || k != 1) dot_partials = nullptr;  // a one-level "multigrid" is the bare smoother apply
    if (!mg->use_graph) return mg_cycle(mg, 0, out, ldo, rhs, ldr, k, true, dot_partials, dot_count);
    GraphKey key{out, rhs, ldo, ldr, k, mg->mu, mg->nu, ctx->option_epoch.load(), dot_partials};
    auto it = mg->graphs.find(key);
    if (it == mg->graphs.end()) {
        if (mg->graphs.size() > 64) {  // bounded cache
            for (auto &g : mg->graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
            mg->graphs.clear();
        }
        const int64_t before = ctx->launches.load();
        CUDA_TRY(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        int ndot = 0;
        famg_status st = mg_cycle(mg, 0, out, ldo, rhs, ldr, k, true, dot_partials, &ndot);
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
        if (st != FAMG_OK) { if (graph) cudaGraphDestroy(graph); return st; }
        if (e != cudaSuccess) FAMG_FAIL(FAMG_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
        GraphEntry ent;
        ent.dot_count = ndot;
        ent.launches = ctx->launches.load() - before;
        ctx->launches.store(before);
        e = cudaGraphInstantiate(&ent.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) FAMG_FAIL(FAMG_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
        it = mg->graphs.emplace(key, ent).first;
    }
    CUDA_TRY(cudaGraphLaunch(it->second.exec, ctx->stream));
    count_launch(ctx, (int)it->second.launches);
    if (dot_count) *dot_count = it->second.dot_count;
    return FAMG_OK;
}

}  // namespace famg

using namespace famg;

extern "C" {

famg_status famg_mg_create(const famg_csr *a0, const famg_smoother *s0, famg_mg **out) {
    if (!a0 || !s0 || !out) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    *out = nullptr;
    if (a0->nrows != a0->ncols || s0->n != a0->nrows) FAMG_FAIL(FAMG_ERR_INVALID, "operator / smoother shape mismatch");
    famg_mg *mg = new famg_mg();
    mg->ctx = a0->ctx;
    MgLevel l;
    l.a = const_cast<famg_csr *>(a0); l.s = const_cast<famg_smoother *>(s0);
    l.a->refs.fetch_add(1); l.s->refs.fetch_add(1);
    mg->lv.push_back(l);
    *out = mg;
    return FAMG_OK;
}

famg_status famg_mg_add_level(famg_mg *mg, const famg_csr *a, const famg_smoother *s, const famg_csr *r, const famg_csr *p) {
    if (!mg || !a || !s || !r || !p) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    const famg_csr *prev = mg->lv.back().a;
    // shape contract of add_level (multigrid.rs:214-227; asserted in hierarchy.rs:258-264)
    if (r->ncols != prev->nrows || p->nrows != prev->nrows || r->nrows != a->nrows || p->ncols != a->nrows ||
        a->nrows != a->ncols || s->n != a->nrows)
        FAMG_FAIL(FAMG_ERR_INVALID, "add_level shape mismatch: prev %lld, A %lldx%lld, R %lldx%lld, P %lldx%lld",
                  (long long)prev->nrows, (long long)a->nrows, (long long)a->ncols, (long long)r->nrows, (long long)r->ncols,
                  (long long)p->nrows, (long long)p->ncols);
    CUDA_TRY(cudaSetDevice(mg->ctx->device));
    cudaStreamSynchronize(mg->ctx->stream);
    mg_free_workspace(mg);
    MgLevel l;
    l.a = const_cast<famg_csr *>(a); l.s = const_cast<famg_smoother *>(s);
    l.r = const_cast<famg_csr *>(r); l.p = const_cast<famg_csr *>(p);
    l.a->refs.fetch_add(1); l.s->refs.fetch_add(1); l.r->refs.fetch_add(1); l.p->refs.fetch_add(1);
    mg->lv.push_back(l);
    return FAMG_OK;
}

famg_status famg_mg_set_cycle(famg_mg *mg, int mu, int smoothing_steps) {
    if (!mg) FAMG_FAIL(FAMG_ERR_INVALID, "null mg");
    if (mu <= 0 || smoothing_steps <= 0) FAMG_FAIL(FAMG_ERR_INVALID, "mu and smoothing_steps must be positive");  // multigrid.rs:204,210
    mg->mu = mu; mg->nu = smoothing_steps;
    return FAMG_OK;
}

famg_status famg_mg_levels(const famg_mg *mg, int *levels) {
    if (!mg || !levels) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    *levels = (int)mg->lv.size();
    return FAMG_OK;
}

famg_status famg_mg_destroy(famg_mg *mg) {
    if (!mg) return FAMG_OK;
    cudaSetDevice(mg->ctx->device);
    cudaStreamSynchronize(mg->ctx->stream);
    mg_free_workspace(mg);
    for (auto &l : mg->lv) {
        csr_release(l.a); smoother_release(l.s);
        if (l.r) csr_release(l.r);
        if (l.p) csr_release(l.p);
    }
    delete mg;
    return FAMG_OK;
}

famg_status famg_mg_apply_dev(famg_mg *mg, famg_vec *out, const famg_vec *rhs) {
    if (!mg || !out || !rhs) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    const int64_t n = mg->lv[0].a->nrows;
    if (out->nrows != n || rhs->nrows != n || out->ncols != rhs->ncols) FAMG_FAIL(FAMG_ERR_INVALID, "multigrid apply shape mismatch");
    if (out->p == rhs->p) FAMG_FAIL(FAMG_ERR_INVALID, "out aliases rhs");
    CUDA_TRY(cudaSetDevice(mg->ctx->device));
    if (rhs->ncols == 0) return FAMG_OK;
    return mg_apply_dev_raw(mg, out->p, out->ld, rhs->p, rhs->ld, (int)rhs->ncols);
}

famg_status famg_mg_apply(famg_mg *mg, double *out, int64_t ld_out, const double *rhs, int64_t ld_rhs, int64_t k) {
    if (!mg || !out || !rhs || k < 0) FAMG_FAIL(FAMG_ERR_INVALID, "bad argument");
    if (k == 0) return FAMG_OK;
    const int64_t n = mg->lv[0].a->nrows;
    famg_vec *x = nullptr, *y = nullptr;
    FAMG_TRY(famg_vec_create(mg->ctx, n, k, &x));
    famg_status st = famg_vec_create(mg->ctx, n, k, &y);
    if (st == FAMG_OK) st = famg_vec_upload(x, rhs, ld_rhs);
    if (st == FAMG_OK) {
        // one-off host-pointer applies do not pollute the graph cache
        const bool g = mg->use_graph; mg->use_graph = false;
        st = famg_mg_apply_dev(mg, y, x);
        mg->use_graph = g;
    }
    if (st == FAMG_OK) st = famg_vec_download(y, out, ld_out);
    famg_vec_destroy(x); famg_vec_destroy(y);
    return st;
}

static double csr_bytes(const famg_csr *m) { return 12.0 * (double)m->nnz + 4.0 * (double)(m->nrows + 1); }

static double cycle_bytes(const famg_mg *mg, size_t level, double k, bool zero_guess) {
    const MgLevel &L = mg->lv[level];
    const double n = (double)L.a->nrows;
    if (level + 1 == mg->lv.size()) {
        if (L.s->kind == SM_DIAG) return 8.0 * n + 16.0 * n * k;
        if (L.s->kind == SM_DENSE_INV) return 8.0 * n * n + 16.0 * n * k;
        return csr_bytes(L.s->minv) + 16.0 * n * k;
    }
    const double sweep = csr_bytes(L.a) + 8.0 * n + 24.0 * n * k;  // read x, b, d; write x'
    const double first = 8.0 * n + 16.0 * n * k;                   // v = d .* f
    double total = (zero_guess ? first + (mg->nu - 1) * sweep : mg->nu * sweep) + mg->nu * sweep;
    total += csr_bytes(L.a) + 24.0 * n * k;                        // residual
    const MgLevel &C = mg->lv[level + 1];
    const double nc = (double)C.a->nrows;
    total += csr_bytes(C.r) + 8.0 * k * (n + nc);                  // restriction
    total += csr_bytes(C.p) + 8.0 * k * (n + nc) + 8.0 * k * n;    // prolongation-add (RMW of v)
    for (int m = 0; m < mg->mu; ++m) total += cycle_bytes(mg, level + 1, k, m == 0);
    return total;
}

famg_status famg_mg_cycle_bytes(const famg_mg *mg, int64_t k, double *bytes) {
    if (!mg || !bytes) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    *bytes = cycle_bytes(mg, 0, (double)k, true);
    return FAMG_OK;
}

}  // extern "C"
