// mg_internal.cuh -- Multigrid internals shared by multigrid.cu and dist.cu.
#pragma once
#include "common.cuh"

struct MgLevel {
    famg_csr *a = nullptr;
    famg_smoother *s = nullptr;
    famg_csr *r = nullptr, *p = nullptr;  // transfer operators to/from the previous (finer) level
    double *x = nullptr, *b = nullptr, *t = nullptr, *z = nullptr;  // workspace, n x ws_k each
    int64_t ld = 0;
};

struct GraphKey {
    const void *out, *rhs; int64_t ldo, ldr; int k, mu, nu;
    int64_t epoch;  // famg_ctx::option_epoch at capture: a changed kernel selection never replays an old graph
    const void *dot;  // partials buffer of the fused rhs . out product (nullptr: none)
    bool operator<(const GraphKey &o) const {
        return std::tie(out, rhs, ldo, ldr, k, mu, nu, epoch, dot) < std::tie(o.out, o.rhs, o.ldo, o.ldr, o.k, o.mu, o.nu, o.epoch, o.dot);
    }
};
struct GraphEntry { cudaGraphExec_t exec = nullptr; int64_t launches = 0; int dot_count = 0; };

struct famg_mg {
    famg_ctx *ctx = nullptr;
    std::vector<MgLevel> lv;
    int mu = 1, nu = 1;
    int ws_k = 0;
    bool use_graph = true;
    std::map<GraphKey, GraphEntry> graphs;
};

namespace famg {
famg_status mg_ensure_workspace(famg_mg *mg, int k);
// one visit of `level`; result in (va, lda); zero_guess: incoming iterate is identically zero.
// dot_partials (k == 1, Diag smoother on this level): the last post-smoothing sweep also leaves the per-CTA
// partial sums of f . va there (*dot_count of them) -- PCG's r.z without another pass over r and z.
famg_status mg_cycle(famg_mg *mg, size_t level, double *va, int64_t lda, const double *f, int64_t ldf, int k, bool zero_guess,
                     double *dot_partials = nullptr, int *dot_count = nullptr);
// out = B rhs through the captured graph; with dot_partials, *dot_count partial sums of rhs . out (0: not fused, the
// caller forms the product itself)
famg_status mg_apply_dev_raw(famg_mg *mg, double *out, int64_t ldo, const double *rhs, int64_t ldr, int k,
                             double *dot_partials = nullptr, int *dot_count = nullptr);
}  // namespace famg
