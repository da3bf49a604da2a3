// famg.hpp -- header-only C++17 mirror of the reference crate's operator API over the C ABI
// (include/famg.h).  The reference is compiled code (Rust) whose toolchain is absent from this
// image; this mirror keeps its type / method names (SparseMatOp, ParSpmmOp, Multigrid, new_l1, ...)
// and error behaviour (reference panics -> famg::Error exceptions) so host code written against the
// crate ports line by line.  Nothing here computes: every call forwards to libfamg.so.
#pragma once
#include <algorithm>
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


// ---------------------------------------------------------------------------------------------
// Device-resident dense block (faer Mat<f64>, column-major).
class DeviceMat {
public:
    DeviceMat(const Context &ctx, int64_t nrows, int64_t ncols = 1) : nrows_(nrows), ncols_(ncols) { check(famg_vec_create(ctx.raw(), nrows, ncols, &h_)); }
    ~DeviceMat() { famg_vec_destroy(h_); }
    DeviceMat(const DeviceMat &) = delete;
    famg_vec *raw() const { return h_; }
    int64_t nrows() const { return nrows_; }
    int64_t ncols() const { return ncols_; }
    void upload(const double *host, int64_t ld) { check(famg_vec_upload(h_, host, ld)); }
    void download(double *host, int64_t ld) const { check(famg_vec_download(h_, host, ld)); }
    void fill(double v) { check(famg_vec_fill(h_, v)); }
private:
    famg_vec *h_ = nullptr;
    int64_t nrows_, ncols_;
};

// partitioners/mod.rs Partition (the *output* of the host-side partitioner): aggregates in CSR
// form, nodes ascending inside each aggregate (BTreeSet order).
struct Partition {
    std::vector<uint64_t> agg_ptr, agg_nodes;
    int64_t naggs() const { return (int64_t)agg_ptr.size() - 1; }
    int64_t nnodes() const { return (int64_t)agg_nodes.size(); }
};

// Deterministic geometric aggregates of a lexicographic nx*ny*nz grid (bx x by x bz boxes; a
// trailing partial box joins its predecessor) -- the benchmark configurations' partitioner.
inline Partition geometric_partition(const int64_t dims[3], const int64_t block[3], int64_t coarse_dims[3]) {
    int64_t c[3];
    for (int d = 0; d < 3; ++d) { c[d] = dims[d] / block[d] > 0 ? dims[d] / block[d] : 1; coarse_dims[d] = c[d]; }
    const int64_t n = dims[0] * dims[1] * dims[2], na = c[0] * c[1] * c[2];
    std::vector<uint64_t> count((size_t)na + 1, 0), agg((size_t)n);
    for (int64_t z = 0; z < dims[2]; ++z)
        for (int64_t y = 0; y < dims[1]; ++y)
            for (int64_t x = 0; x < dims[0]; ++x) {
                const int64_t ax = std::min(x / block[0], c[0] - 1), ay = std::min(y / block[1], c[1] - 1), az = std::min(z / block[2], c[2] - 1);
                const int64_t a = ax + c[0] * (ay + c[1] * az);
                agg[(size_t)(x + dims[0] * (y + dims[1] * z))] = (uint64_t)a;
                ++count[(size_t)a + 1];
            }
    Partition p;
    p.agg_ptr.assign(count.begin(), count.end());
    for (int64_t a = 0; a < na; ++a) p.agg_ptr[(size_t)a + 1] += p.agg_ptr[(size_t)a];
    p.agg_nodes.resize((size_t)n);
    std::vector<uint64_t> fill(p.agg_ptr.begin(), p.agg_ptr.end() - 1);
    for (int64_t i = 0; i < n; ++i) p.agg_nodes[(size_t)fill[(size_t)agg[(size_t)i]]++] = (uint64_t)i;  // ascending node order
    return p;
}

// partitioners/mod.rs PartitionerConfig (:249-329): the algebraic partitioner (least-squares strength
// graph -> greedy modularity matching -> node-swap refinement), host-only, with the documented
// tie-breaking of csrc/partition.cu.  `build_partition` takes the host CSR pattern of the operator.
struct PartitionerConfig {
    double coarsening_factor = 8.0, agg_size_penalty = 1.0;  // :257-266
    int64_t max_improvement_iters = 100;
    Partition build_partition(int64_t n, const uint64_t *row_ptr, const uint64_t *col_idx, const std::vector<double> &near_null, int64_t k,
                              const std::vector<double> &weights) const {
        famg_graph *g = nullptr;
        check(famg_strength_graph_create(n, row_ptr, col_idx, near_null.data(), n, k, weights.data(), 3, &g));  // depth 3, :290
        std::vector<uint64_t> node_to_agg((size_t)n);
        int64_t naggs = 0;
        famg_status st = famg_partition_modularity(g, coarsening_factor, agg_size_penalty, max_improvement_iters, node_to_agg.data(), &naggs);
        famg_graph_destroy(g);
        check(st);
        Partition p;  // Partition::from_node_to_agg (:86-93)
        p.agg_ptr.assign((size_t)naggs + 1, 0);
        for (uint64_t a : node_to_agg) ++p.agg_ptr[(size_t)a + 1];
        for (int64_t a = 0; a < naggs; ++a) p.agg_ptr[(size_t)a + 1] += p.agg_ptr[(size_t)a];
        p.agg_nodes.resize((size_t)n);
        std::vector<uint64_t> fill(p.agg_ptr.begin(), p.agg_ptr.end() - 1);
        for (int64_t i = 0; i < n; ++i) p.agg_nodes[(size_t)fill[(size_t)node_to_agg[(size_t)i]]++] = (uint64_t)i;
        return p;
    }
    Partition build_partition(const SparseRowMat &mat, const std::vector<double> &near_null, int64_t k, const std::vector<double> &weights) const {
        std::vector<uint64_t> rp((size_t)mat.nrows() + 1), ci((size_t)std::max<int64_t>(mat.compute_nnz(), 1));
        std::vector<double> v(ci.size());
        check(famg_csr_download(mat.raw(), rp.data(), ci.data(), v.data()));
        return build_partition(mat.nrows(), rp.data(), ci.data(), near_null, k, weights);
    }
};

// interpolation/mod.rs block_jacobi (:963-1028) and smooth_p (:1030-1040) for block_size > 1
inline std::shared_ptr<SparseRowMat> block_jacobi(const SparseRowMat &mat, int64_t block_size, const SparseRowMat &p) {
    famg_csr *out = nullptr;
    check(famg_block_jacobi(mat.raw(), block_size, p.raw(), &out));
    return std::make_shared<SparseRowMat>(out);
}
inline std::shared_ptr<SparseRowMat> smooth_p(const SparseRowMat &mat, const SparseRowMat &m_inv, const SparseRowMat &p) {
    famg_csr *out = nullptr;
    check(famg_smooth_p(mat.raw(), m_inv.raw(), p.raw(), &out));
    return std::make_shared<SparseRowMat>(out);
}

// block_smoothers.rs BlockSmoother::new (:88-123); vdim > 1 uses diagonally_compensate_vector (:326-400)
inline std::shared_ptr<Smoother> BlockSmoother(const SparseRowMat &mat, int64_t vdim, const Partition &partition) {
    famg_smoother *s = nullptr;
    check(famg_smoother_block_vector(mat.raw(), vdim, partition.naggs(), partition.agg_ptr.data(), partition.agg_nodes.data(), &s));
    return std::make_shared<Smoother>(s);
}

// adaptivity.rs ErrorPropogator (:168-198) and smooth_vector (:307-390), device-resident.
struct ErrorPropogator {
    const SparseRowMat &op;
    const Smoother &pc;
    void apply(DeviceMat &out, const DeviceMat &x) const { check(famg_error_propagator_dev(op.raw(), pc.raw(), out.raw(), x.raw())); }
};
// x: the caller's random start block (n x k, k <= 64), replaced by the smooth basis; returns ||E w||_A / ||w||_A per column
inline std::vector<double> smooth_vector(const SparseRowMat &mat, const Smoother &pc, int64_t iterations, DeviceMat &x) {
    std::vector<double> cfs((size_t)x.ncols());
    check(famg_smooth_vector_dev(mat.raw(), pc.raw(), iterations, x.raw(), cfs.data()));
    return cfs;
}

// preconditioners/composite.rs Composite (:11-100): symmetric multiplicative combination of multigrids around
// one operator; components are kept alive by the shared_ptrs held here.
class Composite {
public:
    Composite(const SparseRowMat &mat, std::shared_ptr<Multigrid> first_component) {
        check(famg_composite_create(mat.raw(), &h_));
        push(std::move(first_component));
    }
    ~Composite() { famg_composite_destroy(h_); }
    Composite(const Composite &) = delete;
    void push(std::shared_ptr<Multigrid> component) {  // composite.rs:85-87
        check(famg_composite_push(h_, FAMG_PC_MG, component->raw()));
        components_.push_back(std::move(component));
    }
    const std::vector<std::shared_ptr<Multigrid>> &components() const { return components_; }
    void apply(DeviceMat &out, const DeviceMat &rhs) const { check(famg_composite_apply_dev(h_, out.raw(), rhs.raw())); }
    famg_composite *raw() const { return h_; }
private:
    famg_composite *h_ = nullptr;
    std::vector<std::shared_ptr<Multigrid>> components_;
};

// interpolation/mod.rs GalerkinCoarse (:34-40)
struct GalerkinCoarse {
    std::shared_ptr<SparseRowMat> interpolation, restriction, coarse_mat;
    std::vector<double> coarse_nn;  // (naggs*cand) x k column-major
    Partition partition;
};

// interpolation/mod.rs smoothed_aggregation (:730-836): tentative P (host SVDs), then
// P = smooth^steps(P0), R = P^T, A_c = R (A P) on the device in one ABI call.
inline GalerkinCoarse smoothed_aggregation(const Context &ctx, const SparseRowMat &fine, Partition partition, int64_t block_size,
                                           const std::vector<double> &near_null, int64_t k, int64_t candidate_dimension,
                                           int smoothing_steps) {
    const int64_t n = fine.nrows();
    GalerkinCoarse g;
    g.coarse_nn.assign((size_t)(partition.naggs() * candidate_dimension * k), 0.0);
    famg_csr *p0 = nullptr, *p = nullptr, *r = nullptr, *ac = nullptr;
    check(famg_tentative_p(ctx.raw(), n, block_size, k, candidate_dimension, near_null.data(), n, partition.naggs(),
                           partition.agg_ptr.data(), partition.agg_nodes.data(), &p0, g.coarse_nn.data()));
    SparseRowMat tentative(p0);
    check(famg_galerkin(fine.raw(), p0, smoothing_steps, 0.66, &p, &r, &ac));  // interpolation/mod.rs:811-828
    g.interpolation = std::make_shared<SparseRowMat>(p);
    g.restriction = std::make_shared<SparseRowMat>(r);
    g.coarse_mat = std::make_shared<SparseRowMat>(ac);
    g.partition = std::move(partition);
    return g;
}

// hierarchy.rs HierarchyConfig / Hierarchy: the coarsen loop (:190-248) over the ABI.  The
// partitioner is a callback (level, fine operator) -> Partition, because the reference's
// PartitionerConfig is host-side graph logic outside the GPU path.
struct HierarchyConfig {
    int64_t coarsest_dim = 1000;          // hierarchy.rs:31
    int64_t max_levels = INT64_MAX;
    int smoothing_steps = 1;              // AggregationConfig, interpolation/mod.rs:72-79
    int64_t candidate_dimension = 1;
};

class Hierarchy {
public:
    template <class Partitioner>
    Hierarchy(const Context &ctx, std::shared_ptr<SparseRowMat> fine, std::vector<double> near_null, int64_t k, const HierarchyConfig &cfg,
              Partitioner &&partitioner)
        : config(cfg) {
        operators.push_back(std::move(fine));
        near_nulls.push_back(std::move(near_null));
        int64_t level = 1, coarse_dim = INT64_MAX;
        while (coarse_dim > cfg.coarsest_dim && level < cfg.max_levels) {                         // hierarchy.rs:199
            const SparseRowMat &a = *operators.back();
            GalerkinCoarse g = smoothed_aggregation(ctx, a, partitioner(level - 1, a), 1, near_nulls.back(), k, cfg.candidate_dimension,
                                                    cfg.smoothing_steps);
            coarse_dim = g.coarse_mat->nrows();
            // coarse near-null: 3-step L1 stationary iteration on the device (:217-226), thin QR (:228)
            auto l1 = new_l1(*g.coarse_mat);
            DeviceMat nn(ctx, coarse_dim, k);
            nn.upload(g.coarse_nn.data(), coarse_dim);
            check(famg_stationary_iteration_dev(g.coarse_mat->raw(), l1->raw(), 3, nn.raw()));
            nn.download(g.coarse_nn.data(), coarse_dim);
            check(famg_thin_q(coarse_dim, k, g.coarse_nn.data(), coarse_dim));
            add_level(std::move(g));
            ++level;
        }
    }
    // hierarchy.rs:250-271 (same shape asserts)
    void add_level(GalerkinCoarse g) {
        if (g.interpolation->nrows() != g.restriction->ncols() || g.interpolation->nrows() != operators.back()->nrows() ||
            g.interpolation->ncols() != g.restriction->nrows() || g.interpolation->ncols() != g.coarse_mat->ncols())
            throw Error(FAMG_ERR_INVALID, "add_level: interpolation / restriction do not match the operators");
        operators.push_back(g.coarse_mat);
        interpolations.push_back(g.interpolation);
        restrictions.push_back(g.restriction);
        partitions.push_back(std::move(g.partition));
        near_nulls.push_back(std::move(g.coarse_nn));
    }
    size_t levels() const { return operators.size(); }
    double op_complexity() const {  // hierarchy.rs:352-360
        double total = 0;
        for (auto &m : operators) total += (double)m->compute_nnz();
        return total / (double)operators[0]->compute_nnz();
    }
    double grid_complexity() const {  // hierarchy.rs:346-350
        double total = 0;
        for (auto &m : operators) total += (double)m->nrows();
        return total / (double)operators[0]->nrows();
    }
    std::vector<std::shared_ptr<SparseRowMat>> operators, interpolations, restrictions;
    std::vector<Partition> partitions;
    std::vector<std::vector<double>> near_nulls;
    HierarchyConfig config;
};

// multigrid.rs MultigridConfig::build (:52-164) with a diagonal smoother on every level but the
// coarsest (exact Cholesky solve), the wiring the reference does by hand in simple_geometric.rs.
struct MultigridConfig {
    int mu = 1, smoothing_steps = 1;
    int smoother_kind = FAMG_DIAG_L1;
    double omega = 0.66;
    std::unique_ptr<Multigrid> build(const Hierarchy &h, std::vector<std::shared_ptr<Smoother>> &keep) const {
        auto make = [&](const SparseRowMat &m, bool coarsest) {
            famg_smoother *s = nullptr;
            if (coarsest) check(famg_smoother_cholesky(m.raw(), &s));
            else check(famg_smoother_diag(m.raw(), smoother_kind, omega, &s));
            keep.push_back(std::make_shared<Smoother>(s));
            return keep.back();
        };
        const size_t nl = h.levels();
        auto s0 = make(*h.operators[0], nl == 1 ? false : false);
        auto mg = std::make_unique<Multigrid>(*h.operators[0], *s0);
        for (size_t l = 1; l < nl; ++l) {
            auto s = make(*h.operators[l], l + 1 == nl);
            mg->add_level(*h.operators[l], *s, *h.restrictions[l - 1], *h.interpolations[l - 1]);
        }
        mg->with_cycle_type(mu).with_smoothing_steps(smoothing_steps);
        return mg;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Hierarchy::coarsen on row slabs (one rank per GPU; the reference is shared-memory only).  A Comm hosts one rank of a
// multi-process job (id from famg_comm_unique_id on rank 0) or -- for tests on a single GPU -- all `nranks` virtual
// ranks of the CONSTRUCTION (Comm::virtual_ranks); per-rank arguments are vectors with one entry per hosted rank.
// ---------------------------------------------------------------------------------------------------------------
class Comm {
public:
    Comm(const Context &ctx, int nranks, int rank, const void *unique_id) { check(famg_comm_create(ctx.raw(), nranks, rank, unique_id, &h_)); }
    static std::unique_ptr<Comm> virtual_ranks(const Context &ctx, int nranks) {
        famg_comm *h = nullptr;
        check(famg_comm_create_sim(ctx.raw(), nranks, &h));
        return std::unique_ptr<Comm>(new Comm(h));
    }
    ~Comm() { famg_comm_destroy(h_); }
    Comm(const Comm &) = delete;
    famg_comm *raw() const { return h_; }
    int nranks() const { int n; check(famg_comm_dims(h_, &n, nullptr, nullptr)); return n; }
    int rank() const { int r; check(famg_comm_dims(h_, nullptr, &r, nullptr)); return r; }
    int nlocal() const { int n; check(famg_comm_dims(h_, nullptr, nullptr, &n)); return n; }
private:
    explicit Comm(famg_comm *h) : h_(h) {}
    famg_comm *h_ = nullptr;
};

// row-partitioned SparseRowMat: the slabs carry global column ids and are taken over (finalize renumbers them in place)
class DistMat {
public:
    DistMat(const Comm &comm, const std::vector<std::shared_ptr<SparseRowMat>> &slabs, int64_t ncols_global) : nlocal_(comm.nlocal()), nranks_(comm.nranks()) {
        std::vector<famg_csr *> raw;
        for (auto &m : slabs) raw.push_back(m->raw());
        check(famg_dmat_create(comm.raw(), raw.data(), ncols_global, nullptr, &h_));
    }
    explicit DistMat(famg_dmat *adopt, int nlocal, int nranks) : h_(adopt), nlocal_(nlocal), nranks_(nranks) {}
    ~DistMat() { famg_dmat_destroy(h_); }
    DistMat(const DistMat &) = delete;
    famg_dmat *raw() const { return h_; }
    DistMat &finalize(bool replicated_cols = false) { check(famg_dmat_finalize(h_, replicated_cols ? 1 : 0)); return *this; }
    int64_t nrows() const { int64_t n; check(famg_dmat_info(h_, &n, nullptr, nullptr, nullptr)); return n; }
    std::vector<int64_t> row_split() const { std::vector<int64_t> rs((size_t)nranks_ + 1); check(famg_dmat_info(h_, nullptr, nullptr, rs.data(), nullptr)); return rs; }
    std::shared_ptr<SparseRowMat> local(int li, bool global_cols) const { famg_csr *m; check(famg_dmat_local(h_, li, global_cols ? 1 : 0, &m)); return std::make_shared<SparseRowMat>(m); }
    std::shared_ptr<SparseRowMat> gather() const {  // the whole operator, replicated (collective); one copy is kept
        std::vector<famg_csr *> out((size_t)nlocal_, nullptr);
        check(famg_dmat_gather(h_, out.data()));
        for (size_t i = 1; i < out.size(); ++i) famg_csr_destroy(out[i]);
        return std::make_shared<SparseRowMat>(out[0]);
    }
private:
    famg_dmat *h_ = nullptr;
    int nlocal_ = 1, nranks_ = 1;
};

// One level: every hosted rank's aggregates over LOCAL row ids and its near-null slice in; P, R, A_c and the coarse near-null
// slices out (hierarchy.rs:203-216 / interpolation/mod.rs:730-836 on slabs).
struct DistGalerkinCoarse {
    std::unique_ptr<DistMat> interpolation, restriction, coarse_mat;
    std::vector<std::vector<double>> coarse_nn;
};
inline DistGalerkinCoarse dist_smoothed_aggregation(const Comm &comm, DistMat &fine, const std::vector<Partition> &parts,
                                                    const std::vector<std::vector<double>> &near_null, int smoothing_steps) {
    const int nl = comm.nlocal();
    std::vector<int64_t> na;
    std::vector<const uint64_t *> ap, an;
    std::vector<const double *> nn;
    DistGalerkinCoarse g;
    std::vector<double *> cnn;
    for (int li = 0; li < nl; ++li) {
        na.push_back(parts[(size_t)li].naggs());
        ap.push_back(parts[(size_t)li].agg_ptr.data()); an.push_back(parts[(size_t)li].agg_nodes.data());
        nn.push_back(near_null[(size_t)li].data());
        g.coarse_nn.emplace_back((size_t)std::max<int64_t>(na.back(), 1), 0.0);
    }
    for (auto &v : g.coarse_nn) cnn.push_back(v.data());
    famg_dmat *p = nullptr, *r = nullptr, *ac = nullptr;
    check(famg_dist_coarsen(fine.raw(), na.data(), ap.data(), an.data(), nn.data(), smoothing_steps, 0.66, &p, &r, &ac, cnn.data()));
    for (int li = 0; li < nl; ++li) g.coarse_nn[(size_t)li].resize((size_t)na[(size_t)li]);
    g.interpolation.reset(new DistMat(p, nl, comm.nranks()));
    g.restriction.reset(new DistMat(r, nl, comm.nranks()));
    g.coarse_mat.reset(new DistMat(ac, nl, comm.nranks()));
    return g;
}
// hierarchy.rs:217-228 on a finalized distributed level
inline void dist_smooth_near_null(DistMat &a, std::vector<std::vector<double>> &nn) {
    std::vector<double *> ptrs;
    for (auto &v : nn) ptrs.push_back(v.data());
    check(famg_dist_smooth_near_null(a.raw(), 3, ptrs.data()));
}

struct CgParams { double abs_tolerance = 0.0, rel_tolerance = 1e-12; int64_t max_iters = 1000; bool zero_guess = true; };
// faer conjugate_gradient as driven by utils.rs:574-609; throws on NoConvergence like Err(CgError).
inline famg_cg_info conjugate_gradient(double *x, const Multigrid &pc, const SparseRowMat &a, const double *b, const CgParams &p) {
    famg_cg_info info{};
    check(famg_pcg_solve(a.raw(), FAMG_PC_MG, pc.raw(), x, b, p.rel_tolerance, p.abs_tolerance, p.max_iters, p.zero_guess, &info));
    return info;
}

}  // namespace famg
