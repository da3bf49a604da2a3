// famg.hpp -- header-only C++17 mirror of the reference crate's operator API over the C ABI
// (include/famg.h).  The reference is compiled code (Rust) whose toolchain is absent from this
// image; this mirror keeps its type / method names (SparseMatOp, ParSpmmOp, Multigrid, new_l1, ...)
// and error behaviour (reference panics -> famg::Error exceptions) so host code written against the
// crate ports line by line.  Nothing here computes: every call forwards to libfamg.so.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "famg.h"

namespace famg {

struct Error : std::runtime_error {
    famg_status status;
    Error(famg_status s, const std::string &m) : std::runtime_error(m), status(s) {}
};
inline void check(famg_status s) {
    if (s != FAMG_OK) throw Error(s, famg_last_error());
}

// One per (process, device): stands where faer::set_global_parallelism(Par::Rayon(n)) stood.
class Context {
public:
    explicit Context(int device = 0) { check(famg_ctx_create(device, &h_)); }
    ~Context() { famg_ctx_destroy(h_); }
    Context(const Context &) = delete;
    famg_ctx *raw() const { return h_; }
    void sync() const { check(famg_ctx_sync(h_)); }
private:
    famg_ctx *h_ = nullptr;
};

// faer SparseRowMat<usize, f64> on the device.
class SparseRowMat {
public:
    SparseRowMat(const Context &ctx, int64_t nrows, int64_t ncols, const uint64_t *row_ptr, const uint64_t *col_idx,
                 const double *val) { check(famg_csr_create(ctx.raw(), nrows, ncols, row_ptr, col_idx, val, &h_)); }
    explicit SparseRowMat(famg_csr *adopt) : h_(adopt) {}
    ~SparseRowMat() { famg_csr_destroy(h_); }
    SparseRowMat(const SparseRowMat &) = delete;
    static std::shared_ptr<SparseRowMat> try_new_from_triplets(const Context &ctx, int64_t nrows, int64_t ncols,
                                                               const std::vector<uint64_t> &r, const std::vector<uint64_t> &c,
                                                               const std::vector<double> &v) {
        famg_csr *h = nullptr;
        check(famg_csr_create_from_triplets(ctx.raw(), nrows, ncols, (int64_t)r.size(), r.data(), c.data(), v.data(), &h));
        return std::make_shared<SparseRowMat>(h);
    }
    famg_csr *raw() const { return h_; }
    int64_t nrows() const { int64_t n; check(famg_csr_dims(h_, &n, nullptr, nullptr)); return n; }
    int64_t ncols() const { int64_t n; check(famg_csr_dims(h_, nullptr, &n, nullptr)); return n; }
    int64_t compute_nnz() const { int64_t n; check(famg_csr_dims(h_, nullptr, nullptr, &n)); return n; }
    // LinOp::apply(out, rhs): column-major host buffers
    void apply(double *out, int64_t ld_out, const double *rhs, int64_t ld_rhs, int64_t k) const {
        check(famg_spmm(h_, out, ld_out, rhs, ld_rhs, k));
    }
    std::shared_ptr<SparseRowMat> transpose() const { famg_csr *t; check(famg_transpose(h_, &t)); return std::make_shared<SparseRowMat>(t); }
    std::shared_ptr<SparseRowMat> operator*(const SparseRowMat &b) const {
        famg_csr *c; check(famg_spgemm(h_, b.raw(), &c)); return std::make_shared<SparseRowMat>(c);
    }
private:
    famg_csr *h_ = nullptr;
};

// par_spmm.rs ParSpmmOp: the accelerated LinOp view of a matrix.
class ParSpmmOp {
public:
    explicit ParSpmmOp(std::shared_ptr<SparseRowMat> m) : mat_(std::move(m)) {}
    int64_t nrows() const { return mat_->nrows(); }
    int64_t ncols() const { return mat_->ncols(); }
    void apply(double *out, int64_t ld_out, const double *rhs, int64_t ld_rhs, int64_t k) const { mat_->apply(out, ld_out, rhs, ld_rhs, k); }
    const std::shared_ptr<SparseRowMat> &mat() const { return mat_; }
private:
    std::shared_ptr<SparseRowMat> mat_;
};

// core.rs SparseMatOp: square CSR + block size + accelerated op (dyn_op()).
class SparseMatOp {
public:
    SparseMatOp(std::shared_ptr<SparseRowMat> m, int64_t block_size = 1) : mat_(std::move(m)), block_size_(block_size) {
        if (mat_->nrows() != mat_->ncols()) throw Error(FAMG_ERR_INVALID, "SparseMatOp is only designed for square sparse matrices.");
        if (mat_->nrows() % block_size != 0) throw Error(FAMG_ERR_INVALID, "Matrix is incompatible with provided block size.");
        par_op_ = std::make_shared<ParSpmmOp>(mat_);
    }
    const std::shared_ptr<SparseRowMat> &arc_mat() const { return mat_; }
    std::shared_ptr<ParSpmmOp> par_op() const { return par_op_; }
    std::shared_ptr<ParSpmmOp> dyn_op() const { return par_op_; }
    int64_t block_size() const { return block_size_; }
private:
    std::shared_ptr<SparseRowMat> mat_;
    std::shared_ptr<ParSpmmOp> par_op_;
    int64_t block_size_;
};

// Arc<dyn BiPrecond<f64>> of one level.
class Smoother {
public:
    explicit Smoother(famg_smoother *adopt) : h_(adopt) {}
    ~Smoother() { famg_smoother_destroy(h_); }
    Smoother(const Smoother &) = delete;
    famg_smoother *raw() const { return h_; }
    void apply(double *out, int64_t ld_out, const double *rhs, int64_t ld_rhs, int64_t k) const {
        check(famg_smoother_apply(h_, out, ld_out, rhs, ld_rhs, k));
    }
private:
    famg_smoother *h_ = nullptr;
};
inline std::shared_ptr<Smoother> new_l1(const SparseRowMat &m) { famg_smoother *s; check(famg_smoother_diag(m.raw(), FAMG_DIAG_L1, 0.0, &s)); return std::make_shared<Smoother>(s); }
inline std::shared_ptr<Smoother> new_l2(const SparseRowMat &m) { famg_smoother *s; check(famg_smoother_diag(m.raw(), FAMG_DIAG_L2, 0.0, &s)); return std::make_shared<Smoother>(s); }
inline std::shared_ptr<Smoother> new_jacobi(const SparseRowMat &m, double omega) { famg_smoother *s; check(famg_smoother_diag(m.raw(), FAMG_DIAG_JACOBI, omega, &s)); return std::make_shared<Smoother>(s); }
inline std::shared_ptr<Smoother> SparseCholeskySolve(const SparseRowMat &m) { famg_smoother *s; check(famg_smoother_cholesky(m.raw(), &s)); return std::make_shared<Smoother>(s); }

// multigrid.rs Multigrid.
class Multigrid {
public:
    Multigrid(const SparseRowMat &op, const Smoother &smoother) { check(famg_mg_create(op.raw(), smoother.raw(), &h_)); }
    ~Multigrid() { famg_mg_destroy(h_); }
    Multigrid(const Multigrid &) = delete;
    void add_level(const SparseRowMat &op, const Smoother &smoother, const SparseRowMat &r, const SparseRowMat &p) {
        check(famg_mg_add_level(h_, op.raw(), smoother.raw(), r.raw(), p.raw()));
    }
    Multigrid &with_cycle_type(int mu) { mu_ = mu; check(famg_mg_set_cycle(h_, mu_, nu_)); return *this; }
    Multigrid &with_smoothing_steps(int nu) { nu_ = nu; check(famg_mg_set_cycle(h_, mu_, nu_)); return *this; }
    int levels() const { int l; check(famg_mg_levels(h_, &l)); return l; }
    void apply(double *out, int64_t ld_out, const double *rhs, int64_t ld_rhs, int64_t k) { check(famg_mg_apply(h_, out, ld_out, rhs, ld_rhs, k)); }
    famg_mg *raw() const { return h_; }
private:
    famg_mg *h_ = nullptr;
    int mu_ = 1, nu_ = 1;
};

struct CgParams { double abs_tolerance = 0.0, rel_tolerance = 1e-12; int64_t max_iters = 1000; bool zero_guess = true; };
// faer conjugate_gradient as driven by utils.rs:574-609; throws on NoConvergence like Err(CgError).
inline famg_cg_info conjugate_gradient(double *x, const Multigrid &pc, const SparseRowMat &a, const double *b, const CgParams &p) {
    famg_cg_info info{};
    check(famg_pcg_solve(a.raw(), FAMG_PC_MG, pc.raw(), x, b, p.rel_tolerance, p.abs_tolerance, p.max_iters, p.zero_guess, &info));
    return info;
}

}  // namespace famg
