// pcg.cu -- device-resident preconditioned conjugate gradients and the stationary solver.
// Reference: faer::matrix_free::conjugate_gradient as driven by test_solver (src/utils.rs:574-609)
// and examples/simple_geometric.rs:229-267; stationary_solver (simple_geometric.rs:117-158).
// faer's source is not in the reference tree; the loop below is the textbook single-rhs PCG with
// faer's stopping rule  ||r||_2 < max(abs_tol, rel_tol * ||b||_2)  (SURVEY 3.3).
//
// Everything stays on the device; the host reads back three scalars per iteration to take the
// convergence decision (one small D2H + stream sync, like the reference's per-iteration norm).
// Kernels per iteration: SpMV with fused p.Ap partials, fused x/r update with ||r||^2 partials,
// the multigrid cycle (one CUDA-graph launch), r.z dot, p update, plus the one-CTA reductions.
#include <cmath>

#include "mg_internal.cuh"

namespace famg {

// z = M^-1 r and r.z into scalar `slot`.  With a multigrid whose fine level has a Diag smoother the product comes
// out of the cycle's last sweep (per-CTA partials, reduced here): no extra pass over r and z.
static famg_status pc_apply_rtz(famg_ctx *ctx, int pc_kind, void *precond, famg_vec *z, const famg_vec *r, int64_t n, int slot) {
    if (pc_kind == FAMG_PC_MG) {
        famg_mg *mg = (famg_mg *)precond;
        if (!z || !r || mg->lv[0].a->nrows != n || z->p == r->p) FAMG_FAIL(FAMG_ERR_INVALID, "multigrid apply shape mismatch");
        FAMG_TRY(ensure_partials(ctx, n + 8));
        int nd = 0;
        FAMG_TRY(mg_apply_dev_raw(mg, z->p, z->ld, r->p, r->ld, 1, ctx->d_partials, &nd));
        if (nd > 0) return reduce_partials(ctx, ctx->d_partials, nd, slot);
        return vec_dot(ctx, r->p, z->p, n, slot);
    }
    FAMG_TRY(pc_apply(pc_kind, precond, z, r));
    return vec_dot(ctx, r->p, z->p, n, slot);
}

famg_status pc_apply(int pc_kind, void *precond, famg_vec *z, const famg_vec *r) {
    if (pc_kind == FAMG_PC_MG) return famg_mg_apply_dev((famg_mg *)precond, z, r);
    if (pc_kind == FAMG_PC_SMOOTHER) return famg_smoother_apply_dev((const famg_smoother *)precond, z, r);
    if (pc_kind == FAMG_PC_COMPOSITE) return famg_composite_apply_dev((famg_composite *)precond, z, r);
    return famg_vec_copy(z, r);
}

}  // namespace famg

using namespace famg;

extern "C" {

// ---- Composite (src/preconditioners/composite.rs) ------------------------------------------------
famg_status famg_composite_create(const famg_csr *a, famg_composite **out) {
    if (!a || !out) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    if (a->nrows != a->ncols) FAMG_FAIL(FAMG_ERR_INVALID, "composite needs a square operator");
    famg_composite *c = new famg_composite();
    c->ctx = a->ctx;
    c->a = const_cast<famg_csr *>(a);
    c->a->refs.fetch_add(1);
    *out = c;
    return FAMG_OK;
}
famg_status famg_composite_push(famg_composite *c, int pc_kind, void *component) {
    if (!c || !component) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    if (pc_kind != FAMG_PC_SMOOTHER && pc_kind != FAMG_PC_MG && pc_kind != FAMG_PC_COMPOSITE) FAMG_FAIL(FAMG_ERR_INVALID, "unknown component kind");
    if (component == (void *)c) FAMG_FAIL(FAMG_ERR_INVALID, "a composite cannot contain itself");
    // components are retained (Arc semantics, like famg_mg's levels): a caller may drop its handle first
    if (pc_kind == FAMG_PC_SMOOTHER) ((famg_smoother *)component)->refs.fetch_add(1);
    c->components.emplace_back(pc_kind, component);
    return FAMG_OK;
}
famg_status famg_composite_len(const famg_composite *c, int64_t *n) {
    if (!c || !n) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    *n = (int64_t)c->components.size();
    return FAMG_OK;
}
famg_status famg_composite_destroy(famg_composite *c) {
    if (!c) return FAMG_OK;
    cudaSetDevice(c->ctx->device);
    for (auto &comp : c->components)
        if (comp.first == FAMG_PC_SMOOTHER) smoother_release((famg_smoother *)comp.second);
    csr_release(c->a);
    delete c;
    return FAMG_OK;
}
// composite.rs:66-83
famg_status famg_composite_apply_dev(famg_composite *c, famg_vec *out, const famg_vec *rhs) {
    if (!c || !out || !rhs) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    const int64_t n = c->a->nrows;
    if (out->nrows != n || rhs->nrows != n || out->ncols != rhs->ncols || out->p == rhs->p) FAMG_FAIL(FAMG_ERR_INVALID, "composite apply shape mismatch");
    famg_ctx *ctx = c->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int k = (int)rhs->ncols;
    famg_vec *ws = nullptr, *tmp = nullptr;
    FAMG_TRY(famg_vec_create(ctx, n, k, &ws));
    famg_status st = famg_vec_create(ctx, n, k, &tmp);
    if (st == FAMG_OK) st = famg_vec_fill(out, 0.0);
    if (st == FAMG_OK) st = famg_vec_copy(ws, rhs);
    auto step = [&](const std::pair<int, void *> &comp) -> famg_status {
        FAMG_TRY(pc_apply(comp.first, comp.second, tmp, ws));                             // ws <- M^-1 ws
        FAMG_TRY(vec_axpby(ctx, out->p, out->ld, tmp->p, tmp->ld, n, k, 1.0, 1.0));       // out += ws
        return famg_residual_dev(c->a, ws, rhs, out);                                     // ws = rhs - A out
    };
    const size_t m = c->components.size();
    for (size_t i = m; i-- > 0 && st == FAMG_OK;) st = step(c->components[i]);
    for (size_t i = 1; i < m && st == FAMG_OK; ++i) st = step(c->components[i]);
    famg_vec_destroy(ws); famg_vec_destroy(tmp);
    return st;
}

famg_status famg_pcg_solve_dev(const famg_csr *a, int pc_kind, void *precond, famg_vec *x, const famg_vec *b, double rel_tol,
                               double abs_tol, int64_t max_iters, int zero_guess, famg_cg_info *info) {
    if (!a || !x || !b || !info) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    if (pc_kind != FAMG_PC_NONE && !precond) FAMG_FAIL(FAMG_ERR_INVALID, "null preconditioner");
    if (a->nrows != a->ncols || x->nrows != a->nrows || b->nrows != a->nrows || x->ncols != 1 || b->ncols != 1)
        FAMG_FAIL(FAMG_ERR_INVALID, "pcg: shape mismatch (single right-hand side required)");
    famg_ctx *ctx = a->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int64_t n = a->nrows;
    info->iter_count = 0; info->abs_residual = 0; info->rel_residual = 0;

    enum { S_BB = 0, S_RR = 1, S_PTQ = 2, S_RTZ_A = 3, S_RTZ_B = 4 };
    double h[8];
    FAMG_TRY(vec_dot(ctx, b->p, b->p, n, S_BB));
    FAMG_TRY(read_scalars(ctx, S_BB, 1, h));
    const double b_norm = sqrt(h[0]);
    if (b_norm == 0.0) return famg_vec_fill(x, 0.0);
    const double thr = std::max(abs_tol, rel_tol * b_norm);

    // work vectors live in a context-owned buffer: stable addresses => the multigrid CUDA graph
    // captured for (z, r) is replayed by every later solve of the same size
    const int64_t ld = (n + 1) & ~(int64_t)1;
    if (ctx->pcg_ws_cap < 4 * ld + 8) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(ctx->pcg_ws); ctx->pcg_ws = nullptr; ctx->pcg_ws_cap = 0;
        FAMG_TRY(dev_alloc(&ctx->pcg_ws, 4 * ld + 8));
        ctx->pcg_ws_cap = 4 * ld + 8;
    }
    famg_vec vr, vp, vz, vq;
    vec_wrap(ctx, ctx->pcg_ws, n, 1, ld, &vr);
    vec_wrap(ctx, ctx->pcg_ws + ld, n, 1, ld, &vp);
    vec_wrap(ctx, ctx->pcg_ws + 2 * ld, n, 1, ld, &vz);
    vec_wrap(ctx, ctx->pcg_ws + 3 * ld, n, 1, ld, &vq);
    famg_vec *r = &vr, *p = &vp, *z = &vz, *q = &vq;
    famg_status st = FAMG_OK;
    double rn = 0.0;
    int slot_rtz = S_RTZ_A, slot_rtz_new = S_RTZ_B;
    bool converged = false;
    do {
        if (st != FAMG_OK) break;
        if (zero_guess) {
            st = famg_vec_fill(x, 0.0);
            if (st == FAMG_OK) st = famg_vec_copy(r, b);
        } else {
            st = famg_residual_dev(a, r, b, x);
        }
        if (st == FAMG_OK) st = vec_dot(ctx, r->p, r->p, n, S_RR);
        if (st == FAMG_OK) st = read_scalars(ctx, S_RR, 1, h);
        if (st != FAMG_OK) break;
        rn = sqrt(h[0]);
        if (rn < thr) { converged = true; break; }
        st = pc_apply_rtz(ctx, pc_kind, precond, z, r, n, slot_rtz);
        if (st == FAMG_OK) st = famg_vec_copy(p, z);
        for (int64_t it = 0; it < max_iters && st == FAMG_OK; ++it) {
            // q = A p, fused partial sums of p.q
            SpmvArgs g; g.a = a; g.epi = EPI_SPMV; g.x = p->p; g.ldx = p->ld; g.y = q->p; g.ldy = q->ld; g.k = 1;
            st = ensure_partials(ctx, n + 8);  // one partial per CTA; never more CTAs than rows
            if (st != FAMG_OK) break;
            g.dot_partials = ctx->d_partials;
            int nct = 0;
            st = spmv_launch(g, &nct);
            if (st == FAMG_OK) st = reduce_partials(ctx, ctx->d_partials, nct, S_PTQ);
            // x += alpha p ; r -= alpha q ; ||r||^2
            if (st == FAMG_OK) st = pcg_update_xr(ctx, x->p, r->p, p->p, q->p, n, slot_rtz, S_PTQ, S_RR);
            if (st == FAMG_OK) st = read_scalars(ctx, S_RR, 4, h);  // rr, ptq, rtz_a, rtz_b
            if (st != FAMG_OK) break;
            const double ptq = h[S_PTQ - S_RR], rtz = h[slot_rtz - S_RR];
            if (!(ptq > 0.0) || !(rtz > 0.0)) {
                set_error("pcg: operator or preconditioner is not positive definite (p.Ap=%g, r.z=%g)", ptq, rtz);
                st = FAMG_ERR_NOT_SPD;
                break;
            }
            rn = sqrt(h[0]);
            info->iter_count = it + 1;
            if (rn < thr) { converged = true; break; }
            st = pc_apply_rtz(ctx, pc_kind, precond, z, r, n, slot_rtz_new);
            if (st == FAMG_OK) st = pcg_update_p(ctx, p->p, z->p, n, slot_rtz_new, slot_rtz);
            std::swap(slot_rtz, slot_rtz_new);
        }
    } while (0);
    info->abs_residual = rn;
    info->rel_residual = rn / b_norm;
    if (st != FAMG_OK) return st;
    if (!converged) {
        set_error("pcg: no convergence in %lld iterations (abs %.3e, rel %.3e)", (long long)max_iters, info->abs_residual,
                  info->rel_residual);
        return FAMG_ERR_NO_CONVERGENCE;
    }
    return FAMG_OK;
}

famg_status famg_pcg_solve(const famg_csr *a, int pc_kind, void *precond, double *x, const double *b, double rel_tol,
                           double abs_tol, int64_t max_iters, int zero_guess, famg_cg_info *info) {
    if (!a || !x || !b || !info) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    famg_vec *dx = nullptr, *db = nullptr;
    FAMG_TRY(famg_vec_create(a->ctx, a->nrows, 1, &dx));
    famg_status st = famg_vec_create(a->ctx, a->nrows, 1, &db);
    if (st == FAMG_OK) st = famg_vec_upload(db, b, a->nrows);
    if (st == FAMG_OK && !zero_guess) st = famg_vec_upload(dx, x, a->nrows);
    famg_status solve = FAMG_OK;
    if (st == FAMG_OK) {
        solve = famg_pcg_solve_dev(a, pc_kind, precond, dx, db, rel_tol, abs_tol, max_iters, zero_guess, info);
        if (solve == FAMG_OK || solve == FAMG_ERR_NO_CONVERGENCE) {
            std::string keep = get_error();
            st = famg_vec_download(dx, x, a->nrows);
            if (solve != FAMG_OK) set_error("%s", keep.c_str());
        }
    }
    famg_vec_destroy(dx); famg_vec_destroy(db);
    return st != FAMG_OK ? st : solve;
}

famg_status famg_stationary_solve(const famg_csr *a, int pc_kind, void *precond, double *x, const double *b, double rel_tol,
                                  int64_t max_iters, int64_t *iters) {
    if (!a || !x || !b || !iters) FAMG_FAIL(FAMG_ERR_INVALID, "null argument");
    if (a->nrows != a->ncols) FAMG_FAIL(FAMG_ERR_INVALID, "square operator required");
    famg_ctx *ctx = a->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int64_t n = a->nrows;
    famg_vec *dx = nullptr, *db = nullptr, *r = nullptr, *z = nullptr;
    famg_status st = famg_vec_create(ctx, n, 1, &dx);
    if (st == FAMG_OK) st = famg_vec_create(ctx, n, 1, &db);
    if (st == FAMG_OK) st = famg_vec_create(ctx, n, 1, &r);
    if (st == FAMG_OK) st = famg_vec_create(ctx, n, 1, &z);
    if (st == FAMG_OK) st = famg_vec_upload(db, b, n);
    if (st == FAMG_OK) st = famg_vec_upload(dx, x, n);
    double h[2];
    if (st == FAMG_OK) st = vec_dot(ctx, db->p, db->p, n, 0);
    if (st == FAMG_OK) st = read_scalars(ctx, 0, 1, h);
    const double b_norm = st == FAMG_OK ? sqrt(h[0]) : 1.0;
    int64_t iter = 0;
    while (st == FAMG_OK) {
        st = famg_residual_dev(a, r, db, dx);
        if (st == FAMG_OK) st = vec_dot(ctx, r->p, r->p, n, 1);
        if (st == FAMG_OK) st = read_scalars(ctx, 1, 1, h);
        if (st != FAMG_OK) break;
        const double rel = sqrt(h[0]) / b_norm;
        ++iter;
        if (rel < rel_tol || iter >= max_iters) break;
        st = pc_apply(pc_kind, precond, z, r);
        if (st == FAMG_OK) st = vec_add_inplace(ctx, dx->p, z->p, n);
    }
    *iters = iter;
    if (st == FAMG_OK) st = famg_vec_download(dx, x, n);
    famg_vec_destroy(dx); famg_vec_destroy(db); famg_vec_destroy(r); famg_vec_destroy(z);
    return st;
}

}  // extern "C"
