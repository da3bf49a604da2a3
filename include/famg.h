/*
 * famg.h -- C ABI of the B200-native faer-amg hot path (libfamg.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types, no exceptions
 * across it.  Every entry point names the interface of the reference crate (aujxn/faer-amg,
 * paths relative to the crate root) that it replaces.  The Rust-side binding a maintainer would
 * add is shown in INTEGRATION.md (rust/faer-amg-b200/src/ffi.rs).
 *
 * Conventions
 *   - Host CSR inputs use the reference's index width: `usize` == uint64_t row_ptr[nrows+1],
 *     col_idx[nnz] (sorted, unique per row, compressed), f64 values.  On the device indices are
 *     32-bit (nnz and ncols must be < 2^31 per device matrix).
 *   - Dense multivectors are column-major with unit row stride (faer Mat<f64>): ptr, ld >= nrows.
 *   - Every function returns famg_status (0 = OK).  On failure famg_last_error() returns a
 *     message for the calling thread; the Rust shim turns non-zero into panic! to keep the
 *     reference's error convention (par_spmm.rs:35,78; hierarchy.rs:259-264).
 *   - Handles are immutable after creation and reference-counted internally: a multigrid keeps
 *     its level operators alive after the caller destroys its own handles (Arc semantics,
 *     core.rs:11-12, multigrid.rs:172-179).
 *   - All work of one famg_ctx is ordered on that context's CUDA stream.  Functions taking host
 *     pointers stage through device memory and return after the result is on the host;
 *     `_dev` functions only enqueue.
 *   - There is no CPU fallback: every entry point that computes fails with FAMG_ERR_CUDA when
 *     no sm_100 device is usable.
 */
#ifndef FAMG_H
#define FAMG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int famg_status;
enum {
    FAMG_OK = 0,
    FAMG_ERR_INVALID = 1,      /* bad argument / shape mismatch (reference: assert!/panic!) */
    FAMG_ERR_CUDA = 2,         /* CUDA runtime failure, no device */
    FAMG_ERR_ALLOC = 3,
    FAMG_ERR_NUMERIC = 4,      /* missing/non-positive diagonal, Cholesky breakdown */
    FAMG_ERR_NO_CONVERGENCE = 5, /* CgError::NoConvergence (utils.rs:649-653) */
    FAMG_ERR_NOT_SPD = 6,      /* CgError non-positive-definite variants */
    FAMG_ERR_COMM = 7,         /* NCCL failure */
    FAMG_ERR_UNSUPPORTED = 8
};

typedef struct famg_ctx famg_ctx;
typedef struct famg_csr famg_csr;           /* device CSR: SparseRowMat<usize,f64> (core.rs:14) */
typedef struct famg_vec famg_vec;           /* device Mat<f64>, column-major */
typedef struct famg_smoother famg_smoother; /* Arc<dyn BiPrecond<f64>> of one level */
typedef struct famg_mg famg_mg;             /* Multigrid (multigrid.rs:171-179) */
typedef struct famg_comm famg_comm;         /* one rank of a row-partitioned multi-GPU job */
typedef struct famg_dist_mg famg_dist_mg;   /* row-partitioned Multigrid + PCG */
typedef struct famg_partition famg_partition; /* device-resident aggregates (Partition, partitioners/mod.rs:23-27) */

const char *famg_last_error(void);
const char *famg_version(void);

/* ---- context ---------------------------------------------------------------------------- */
/* One context per (process, device).  Replaces faer::set_global_parallelism / Par::Rayon(n)
 * (examples/amg/main.rs:227-228) as the "where does the work run" switch. */
famg_status famg_ctx_create(int device, famg_ctx **out);
famg_status famg_ctx_destroy(famg_ctx *ctx);
famg_status famg_ctx_sync(famg_ctx *ctx);
/* the context's cudaStream_t (as void*) so a caller can order its own work / events on it */
famg_status famg_ctx_stream(famg_ctx *ctx, void **stream);
/* SM count, device name, free/total memory -- for reports */
famg_status famg_ctx_info(famg_ctx *ctx, int *num_sms, int64_t *mem_free, int64_t *mem_total,
                          char *name, int name_len);

/* ---- CSR operator: SparseMatOp::new / ParSpmmOp::new (core.rs:56-74, par_spmm.rs:31-96) ---- */
famg_status famg_csr_create(famg_ctx *ctx, int64_t nrows, int64_t ncols, const uint64_t *row_ptr,
                            const uint64_t *col_idx, const double *val, famg_csr **out);
/* SparseRowMat::try_new_from_triplets (interpolation/mod.rs:807, simple_geometric.rs:76):
 * sorted unique columns per row, duplicates summed in input order, explicit zeros kept. */
famg_status famg_csr_create_from_triplets(famg_ctx *ctx, int64_t nrows, int64_t ncols, int64_t nt,
                                          const uint64_t *rows, const uint64_t *cols,
                                          const double *vals, famg_csr **out);
famg_status famg_csr_retain(famg_csr *a);
famg_status famg_csr_destroy(famg_csr *a); /* release */
famg_status famg_csr_dims(const famg_csr *a, int64_t *nrows, int64_t *ncols, int64_t *nnz);
/* caller-allocated host buffers (row_ptr: nrows+1, col_idx/val: nnz) so Rust owns what it wraps
 * in SparseRowMat::new(SymbolicSparseRowMat::new_checked(..), vals). */
famg_status famg_csr_download(const famg_csr *a, uint64_t *row_ptr, uint64_t *col_idx, double *val);
/* row-length statistics that drive the kernel selection (threads-per-row, staged fraction) */
famg_status famg_csr_plan(const famg_csr *a, int *threads_per_row, int *rows_per_cta,
                          double *avg_row_nnz, int *max_row_nnz);
/* rows [row_begin,row_end) of `a` as a new matrix with the same column space
 * (row partitioning for multi-GPU) */
famg_status famg_csr_row_slab(const famg_csr *a, int64_t row_begin, int64_t row_end, famg_csr **out);

/* Synthetic operators generated on the device (the reference ships no 3-D
 * generator, F8).  G7: 7-point Laplacian; G27: 27-point anisotropic diffusion. */
famg_status famg_gallery_g7(famg_ctx *ctx, int64_t nx, int64_t ny, int64_t nz, famg_csr **out);
famg_status famg_gallery_g27(famg_ctx *ctx, int64_t nx, int64_t ny, int64_t nz, double eps_y,
                             double eps_z, famg_csr **out);

/* rows of the planes [z0, z1) of the same operators, with global column ids: one rank's slab of a
 * row-partitioned operator generated in place (nothing global is ever materialised) */
famg_status famg_gallery_g7_slab(famg_ctx *ctx, int64_t nx, int64_t ny, int64_t nz, int64_t z0, int64_t z1, famg_csr **out);
famg_status famg_gallery_g27_slab(famg_ctx *ctx, int64_t nx, int64_t ny, int64_t nz, double eps_y, double eps_z, int64_t z0,
                                  int64_t z1, famg_csr **out);

/* ---- dense multivectors on the device ----------------------------------------------------- */
famg_status famg_vec_create(famg_ctx *ctx, int64_t nrows, int64_t ncols, famg_vec **out);
famg_status famg_vec_destroy(famg_vec *v);
famg_status famg_vec_dims(const famg_vec *v, int64_t *nrows, int64_t *ncols);
famg_status famg_vec_upload(famg_vec *v, const double *host, int64_t ld);
famg_status famg_vec_download(const famg_vec *v, double *host, int64_t ld);
famg_status famg_vec_fill(famg_vec *v, double value);
famg_status famg_vec_copy(famg_vec *dst, const famg_vec *src);
/* y = alpha * x + beta * y (column by column; the vector algebra of ErrorPropogator::apply,
 * adaptivity.rs:191-198, and StationaryIteration::transpose_apply, smoothers.rs:179-197) */
famg_status famg_vec_axpby(famg_vec *y, double alpha, const famg_vec *x, double beta);
/* raw device pointer + leading dimension (for CUDA-event timing harnesses / torch interop) */
famg_status famg_vec_ptr(const famg_vec *v, void **dev_ptr, int64_t *ld);
/* column-wise ||v_j||_2 into host `out` (ncols doubles) */
famg_status famg_vec_norm2(const famg_vec *v, double *out);

/* ---- mat-apply: ParSpmmOp::apply / LinOp::apply (par_spmm.rs:98-158) ----------------------- */
/* out = A * rhs with host buffers (the literal LinOp::apply(out: MatMut, rhs: MatRef) drop-in) */
famg_status famg_spmm(const famg_csr *a, double *out, int64_t ld_out, const double *rhs,
                      int64_t ld_rhs, int64_t k);
famg_status famg_spmm_dev(const famg_csr *a, famg_vec *out, const famg_vec *rhs);
/* out = rhs_b - A * x   (multigrid.rs:341-342, fused) */
famg_status famg_residual_dev(const famg_csr *a, famg_vec *out, const famg_vec *b, const famg_vec *x);
/* y += A * x            (multigrid.rs:349-350, fused prolongation-correct) */
famg_status famg_spmm_add_dev(const famg_csr *a, famg_vec *y, const famg_vec *x);

/* ---- smoothers ---------------------------------------------------------------------------- */
enum { FAMG_DIAG_L1 = 0, FAMG_DIAG_L2 = 1, FAMG_DIAG_JACOBI = 2 };
/* new_l1 / new_l2 / new_jacobi (smoothers.rs:63-76, 43-61, 78-86) -> faer Diag<f64> */
famg_status famg_smoother_diag(const famg_csr *a, int kind, double omega, famg_smoother **out);
/* a caller-computed Diag (any LinOp that is a row scaling) */
famg_status famg_smoother_diag_from_host(famg_ctx *ctx, int64_t n, const double *d, famg_smoother **out);
/* CoarseSolverKind::Cholesky.build_from_sparse -> SparseCholeskySolve (coarse_solvers.rs:20-31,
 * 173-206): exact solve; realised as host Cholesky + explicit inverse applied as a device GEMV */
famg_status famg_smoother_cholesky(const famg_csr *a, famg_smoother **out);
/* BlockSmoother::new over aggregates with diagonal compensation (block_smoothers.rs:88-123,
 * 293-324), materialised as the block-diagonal M^-1 of into_sparse_mat (:125-146).
 * agg_ptr[n_aggs+1], agg_nodes ascending inside each aggregate.  The blocks are assembled, Cholesky-factorised and
 * inverted on the device (aggregates of at most 64 nodes of similar size; otherwise, or with
 * famg_ctx_set_option(ctx, "block_build_host", 1), by the host routine -- same statements, same bits). */
famg_status famg_smoother_block(const famg_csr *a, int64_t n_aggs, const uint64_t *agg_ptr,
                                const uint64_t *agg_nodes, famg_smoother **out);
/* the same for vector problems: the partition is over nodes of `vdim` dofs (dof = node * vdim + offset);
 * off-aggregate coupling blocks are lumped by diagonally_compensate_vector (block_smoothers.rs:326-400) */
famg_status famg_smoother_block_vector(const famg_csr *a, int64_t vdim, int64_t n_aggs, const uint64_t *agg_ptr,
                                       const uint64_t *agg_nodes, famg_smoother **out);
famg_status famg_smoother_retain(famg_smoother *s);
famg_status famg_smoother_destroy(famg_smoother *s);
famg_status famg_smoother_dim(const famg_smoother *s, int64_t *n);
/* Diag smoothers only: copy d to host (n doubles) */
famg_status famg_smoother_diag_download(const famg_smoother *s, double *d);
/* LinOp::apply: out = M^-1 rhs (host buffers / device vectors) */
famg_status famg_smoother_apply(const famg_smoother *s, double *out, int64_t ld_out,
                                const double *rhs, int64_t ld_rhs, int64_t k);
famg_status famg_smoother_apply_dev(const famg_smoother *s, famg_vec *out, const famg_vec *rhs);
/* smooth() (multigrid.rs:407-424): iters x { x += M^-1 (b - A x) }, one fused kernel per sweep
 * for Diag smoothers (A is read once per sweep). */
famg_status famg_smooth_dev(const famg_csr *a, const famg_smoother *s, famg_vec *x, const famg_vec *b,
                            int iters);
/* StationaryIteration::apply_in_place with a Diag preconditioner (smoothers.rs:146-159 as used
 * by hierarchy.rs:219-226), restated literally including r = x - A x. */
famg_status famg_stationary_iteration_dev(const famg_csr *a, const famg_smoother *diag, int iters,
                                          famg_vec *io);

/* ---- Multigrid (multigrid.rs:190-249, 251-380, 469-473) ------------------------------------ */
famg_status famg_mg_create(const famg_csr *a0, const famg_smoother *s0, famg_mg **out); /* new */
famg_status famg_mg_add_level(famg_mg *mg, const famg_csr *a, const famg_smoother *s,
                              const famg_csr *r, const famg_csr *p);                /* add_level */
famg_status famg_mg_set_cycle(famg_mg *mg, int mu, int smoothing_steps); /* with_cycle_type/_steps */
famg_status famg_mg_levels(const famg_mg *mg, int *levels);
famg_status famg_mg_destroy(famg_mg *mg);
/* Multigrid::apply: out = B rhs, one mu-cycle from a zero guess; host buffers */
famg_status famg_mg_apply(famg_mg *mg, double *out, int64_t ld_out, const double *rhs,
                          int64_t ld_rhs, int64_t k);
famg_status famg_mg_apply_dev(famg_mg *mg, famg_vec *out, const famg_vec *rhs);
/* algorithmic HBM bytes of one cycle with k right-hand sides (SURVEY 8(d) formulas) */
famg_status famg_mg_cycle_bytes(const famg_mg *mg, int64_t k, double *bytes);

/* ---- hierarchy construction: the sparse expressions of smoothed_aggregation ---------------- */
/* `&a * &b` on SparseRowMat (interpolation/mod.rs:828,938): sorted, unpruned structural product;
 * every entry accumulates in ascending inner index. */
famg_status famg_spgemm(const famg_csr *a, const famg_csr *b, famg_csr **out);
/* p.transpose().to_row_major() (interpolation/mod.rs:824-827) */
famg_status famg_transpose(const famg_csr *a, famg_csr **out);
/* smooth_interpolation(mat, p, jacobi_weight) (interpolation/mod.rs:927-946), fused into the
 * SpGEMM numeric epilogue */
famg_status famg_smooth_interpolation(const famg_csr *a, const famg_csr *p, double omega,
                                      famg_csr **out);
/* interpolation/mod.rs:811-828 in one call: P = smooth^steps(P0); R = P^T; A_c = R (A P). */
famg_status famg_galerkin(const famg_csr *a, const famg_csr *p0, int smoothing_steps, double omega,
                          famg_csr **p, famg_csr **r, famg_csr **a_coarse);
/* the same for block_size > 1: prolongator smoothing by block_jacobi (interpolation/mod.rs:816, 963-1028)
 * -- D^-1 = -0.66 * inverse of A's block diagonal (per-block symmetric eigen-decomposition of the lower
 * side, eigenvalues must exceed 1e-6), P <- D^-1 (A P) + P -- instead of smooth_interpolation.  omega is
 * used only when block_size == 1. */
famg_status famg_galerkin_block(const famg_csr *a, const famg_csr *p0, int64_t block_size, int smoothing_steps,
                                double omega, famg_csr **p, famg_csr **r, famg_csr **a_coarse);
/* block_jacobi (interpolation/mod.rs:963-1028) and smooth_p (:1030-1040: m_inv * (-(A P)) + P) on their own */
famg_status famg_block_jacobi(const famg_csr *a, int64_t block_size, const famg_csr *p, famg_csr **out);
famg_status famg_smooth_p(const famg_csr *a, const famg_csr *m_inv, const famg_csr *p, famg_csr **out);
/* Deterministic geometric aggregates of a lexicographic grid (the benchmark configurations' stand-in for the
 * caller's partitioner): bx x by x bz boxes, trailing partial boxes join their predecessor, nodes
 * ascending per aggregate.  Host-only.  agg_ptr: prod(coarse_dims) + 1 entries, agg_nodes: nx*ny*nz; pass NULL
 * for both to query coarse_dims[3] first. */
famg_status famg_geometric_partition(int64_t nx, int64_t ny, int64_t nz, int64_t bx, int64_t by, int64_t bz,
                                     uint64_t *agg_ptr, uint64_t *agg_nodes, int64_t *coarse_dims);
/* tentative prolongator of smoothed_aggregation (interpolation/mod.rs:747-809): per-aggregate
 * thin SVD of the near-null block on the host (tiny dense work), P uploaded as CSR.
 * near_null: n_fine x k column-major; coarse_nn (out): (n_aggs*cand) x k column-major. */
famg_status famg_tentative_p(famg_ctx *ctx, int64_t n_fine, int64_t block_size, int64_t k,
                             int64_t cand, const double *near_null, int64_t ld_nn, int64_t n_aggs,
                             const uint64_t *agg_ptr, const uint64_t *agg_nodes, famg_csr **p,
                             double *coarse_nn);
/* The same two steps with the aggregates in device memory (the north-star moves aggregation work to the device where it
 * is measured on the critical path: these two host passes were a quarter of a 256^3 hierarchy build, and grew with the
 * number of ranks sharing the host cores).  famg_partition: aggregate lists on the device -- generated in place for box
 * aggregates (same aggregates and order as famg_geometric_partition) or uploaded from any host partitioner. */
famg_status famg_partition_geometric_dev(famg_ctx *ctx, int64_t nx, int64_t ny, int64_t nz, int64_t bx, int64_t by, int64_t bz,
                                         famg_partition **out, int64_t *coarse_dims /* 3, may be NULL */);
famg_status famg_partition_upload(famg_ctx *ctx, int64_t n_nodes, int64_t n_aggs, const uint64_t *agg_ptr, const uint64_t *agg_nodes,
                                  famg_partition **out);
famg_status famg_partition_dims(const famg_partition *p, int64_t *n_nodes, int64_t *n_aggs);
famg_status famg_partition_download(const famg_partition *p, uint64_t *agg_ptr, uint64_t *agg_nodes);
famg_status famg_partition_destroy(famg_partition *p);
/* tentative prolongator (interpolation/mod.rs:747-809) of a scalar problem with one near-null vector, on the device: one
 * thread per aggregate in list order -- bit-identical to famg_tentative_p.  near_null: n_nodes x 1, coarse_nn: n_aggs x 1 */
famg_status famg_tentative_p_dev(const famg_partition *part, const famg_vec *near_null, famg_csr **p, famg_vec *coarse_nn);
/* coarse_nn.qr().compute_thin_Q() (hierarchy.rs:228); host, in place, R with positive diagonal */
famg_status famg_thin_q(int64_t n, int64_t k, double *a, int64_t lda);
/* the same on a device-resident block (k <= 64): CholeskyQR2 -- Gram matrix by a deterministic
 * reduction, k x k Cholesky on the host, X <- X R^-1, twice.  Keeps the near-null search loop
 * (adaptivity.rs:331-354: one thin QR per error-propagation step on an n x k block) in HBM. */
famg_status famg_thin_q_dev(famg_vec *x);

/* ---- near-null search kept on the device -------------------------------------- */
/* ErrorPropogator::apply (adaptivity.rs:191-198): out = x - M^-1 (A x) on an n x k block.  With a
 * Diag preconditioner this is one fused SpMM launch (A and x read once). */
famg_status famg_error_propagator_dev(const famg_csr *a, const famg_smoother *s, famg_vec *out, const famg_vec *x);
/* smooth_vector (adaptivity.rs:307-390) from the caller's random block x (n x k, k <= 64), in place:
 * x = thinQ(thinQ(x)); iterations x { x = E x; x = thinQ(x) }   (:331, :346, :351-354)
 * cfs (k doubles, may be NULL): ||E w||_A / ||w||_A per column (:365-384).  The reference draws x
 * from an unseeded StandardNormal stream (:321-329); the caller supplies it here. */
famg_status famg_smooth_vector_dev(const famg_csr *a, const famg_smoother *s, int64_t iterations, famg_vec *x, double *cfs);
/* out[c] = x[:,c] . y[:,c] (deterministic two-stage reduction) */
famg_status famg_vec_coldot(const famg_vec *x, const famg_vec *y, double *out);

/* ---- Composite preconditioner (src/preconditioners/composite.rs) -------------------------------- */
/* Symmetric multiplicative combination of several preconditioners around one operator (composite.rs:66-83):
 *   out = 0; ws = rhs; for c in components reversed, then components[1..]: ws <- c^-1 ws; out += ws; ws = rhs - A out
 * Device-resident: the `rhs - A out` step is the fused residual kernel.  Components are (pc_kind, handle)
 * pairs: smoothers are retained by the composite; multigrid / composite components are borrowed and must
 * outlive it (the Rust side holds the Arcs).
 * Usable wherever a (pc_kind, precond) pair is accepted: famg_pcg_solve*, famg_stationary_solve,
 * famg_smooth_vector_pc_dev (the adaptive driver's test loop, adaptivity.rs:108-114). */
typedef struct famg_composite famg_composite;
famg_status famg_composite_create(const famg_csr *a, famg_composite **out);      /* Composite::new_with_components(mat, vec![]) */
famg_status famg_composite_push(famg_composite *c, int pc_kind, void *component); /* composite.rs:85-87 */
famg_status famg_composite_len(const famg_composite *c, int64_t *n);
famg_status famg_composite_apply_dev(famg_composite *c, famg_vec *out, const famg_vec *rhs);
famg_status famg_composite_destroy(famg_composite *c);
/* smooth_vector (adaptivity.rs:307-390) with any preconditioner kind; famg_smooth_vector_dev is the FAMG_PC_SMOOTHER case */
famg_status famg_smooth_vector_pc_dev(const famg_csr *a, int pc_kind, void *precond, int64_t iterations, famg_vec *x,
                                      double *cfs);

/* ---- host partitioner: PartitionerConfig::build_partition, partitioners/mod.rs:273-329 -
 * Host-only (no device work, no ctx): the north-star keeps aggregation on the host.  A Rust build keeps
 * using the crate's own partitioner; these entry points give the non-Rust harness algebraic aggregates.
 * Tie-breaking where the reference is unspecified is documented in csrc/partition.cu. */
typedef struct famg_graph famg_graph; /* AdjacencyList (partitioners/mod.rs:331-334) */
/* AdjacencyList::new_ls_strength_graph (partitioners/mod.rs:337-393); the reference uses max_depth 3 (:290).
 * CSR pattern of the square matrix (values are not used), near_null n x k column-major, k weights. */
famg_status famg_strength_graph_create(int64_t n, const uint64_t *row_ptr, const uint64_t *col_idx,
                                       const double *near_null, int64_t ldn, int64_t k, const double *weights,
                                       int64_t max_depth, famg_graph **out);
/* an explicit weighted adjacency list (PartitionerConfig::build_from_strength, partitioners/mod.rs:310-317) */
famg_status famg_graph_create(int64_t n, const uint64_t *row_ptr, const uint64_t *col_idx, const double *w,
                              famg_graph **out);
/* block_size > 1: strength.aggregate(&block_reduce) + filter_diag (partitioners/mod.rs:293-300), in place:
 * the graph over dofs becomes a graph over nodes of block_size consecutive dofs */
famg_status famg_graph_block_reduce(famg_graph *g, int64_t block_size);
famg_status famg_graph_dims(const famg_graph *g, int64_t *n, int64_t *nnz);
famg_status famg_graph_download(const famg_graph *g, uint64_t *row_ptr, uint64_t *col_idx, double *w);
famg_status famg_graph_destroy(famg_graph *g);
/* Partitioner::new + initialize_partition + improve_partition (partitioners/modularity.rs:28-137, 179-192,
 * 437-510) from the singleton partition; node_to_agg has n entries, aggregates are numbered 0..naggs-1. */
famg_status famg_partition_modularity(const famg_graph *g, double coarsening_factor, double agg_size_penalty,
                                      int64_t max_improvement_iters, uint64_t *node_to_agg, int64_t *naggs);

/* ---- PCG: faer conjugate_gradient as driven by utils.rs:574-609 ----------------------------- */
typedef struct {
    int64_t iter_count;
    double abs_residual;
    double rel_residual;
} famg_cg_info;
enum { FAMG_PC_NONE = 0, FAMG_PC_SMOOTHER = 1, FAMG_PC_MG = 2, FAMG_PC_COMPOSITE = 3 };
/* x (in: initial guess unless zero_guess; out: solution) and b are host buffers of a.nrows
 * doubles.  precond: famg_smoother* or famg_mg* according to pc_kind.  Returns
 * FAMG_ERR_NO_CONVERGENCE with info filled when max_iters is hit. */
famg_status famg_pcg_solve(const famg_csr *a, int pc_kind, void *precond, double *x, const double *b,
                           double rel_tol, double abs_tol, int64_t max_iters, int zero_guess,
                           famg_cg_info *info);
famg_status famg_pcg_solve_dev(const famg_csr *a, int pc_kind, void *precond, famg_vec *x,
                               const famg_vec *b, double rel_tol, double abs_tol, int64_t max_iters,
                               int zero_guess, famg_cg_info *info);
/* stationary_solver (examples/simple_geometric.rs:117-158); returns iterations in *iters */
famg_status famg_stationary_solve(const famg_csr *a, int pc_kind, void *precond, double *x,
                                  const double *b, double rel_tol, int64_t max_iters, int64_t *iters);

/* ---- multi-GPU: 1-D row partition.  Data path: peer-memory (CUDA IPC over NVLink) halo
 * exchange and small collectives; NCCL bootstraps the job, carries the setup-time exchanges and is the
 * fallback data path when a peer cannot be mapped. ------------------------------------------------- */
#define FAMG_UNIQUE_ID_BYTES 128
famg_status famg_comm_unique_id(void *id_bytes /* 128 */);
famg_status famg_comm_create(famg_ctx *ctx, int nranks, int rank, const void *id_bytes, famg_comm **out);
famg_status famg_comm_destroy(famg_comm *c);
/* sum-all-reduce of n doubles on the host (setup-time plumbing; also a smoke test of the comm) */
famg_status famg_comm_allreduce_sum(famg_comm *c, double *vals, int n);
/* Row-partitioned multigrid from a replicated global hierarchy: level l owns rows
 * [row_splits[l][rank], row_splits[l][rank+1]).  Levels with fewer than `replicate_below` rows
 * per rank are run replicated after an all-gather of the restricted residual. */
famg_status famg_dist_mg_create(famg_comm *c, famg_mg *global_mg, const int64_t *const *row_splits,
                                int64_t replicate_below, famg_dist_mg **out);
famg_status famg_dist_mg_destroy(famg_dist_mg *d);
/* PCG on the partitioned system: x_local/b_local are this rank's rows (host buffers). */
famg_status famg_dist_pcg_solve(famg_dist_mg *d, double *x_local, const double *b_local,
                                double rel_tol, double abs_tol, int64_t max_iters, int zero_guess,
                                famg_cg_info *info);
famg_status famg_dist_pcg_solve_dev(famg_dist_mg *d, famg_vec *x_local, const famg_vec *b_local,
                                    double rel_tol, double abs_tol, int64_t max_iters,
                                    int zero_guess, famg_cg_info *info);
/* distributed y_local = A x_local (halo exchange + overlap), device vectors of local rows */
famg_status famg_dist_spmv_dev(famg_dist_mg *d, famg_vec *y_local, const famg_vec *x_local);
famg_status famg_dist_mg_apply_dev(famg_dist_mg *d, famg_vec *out_local, const famg_vec *rhs_local);

/* ---- distributed hierarchy construction (SURVEY 8(e) "RAP"; Hierarchy::coarsen, hierarchy.rs:190-248, on row
 * slabs).  Every rank builds only its rows of P = (I - w D^-1 A) P0, A P and A_c = R (A P), fetching the off-rank
 * rows of P0 / P / AP it needs; nothing global is materialised.  A famg_dmat is a row-partitioned sparse matrix:
 * one slab per virtual rank hosted by the calling process -- one with a process per GPU; all `nranks` of them for
 * a communicator made by famg_comm_create_sim, which runs the same construction inside one process on one GPU
 * (exchanges become device copies; setup only, it cannot run the distributed cycle).  Per-rank array arguments
 * below have one entry per hosted virtual rank (famg_comm_dims: nlocal). */
typedef struct famg_dmat famg_dmat;
famg_status famg_comm_create_sim(famg_ctx *ctx, int nranks, famg_comm **out);
famg_status famg_comm_dims(const famg_comm *c, int *nranks, int *rank, int *nlocal);
/* host all-gather of per-rank pieces (rank order): out[li] receives sum(counts) doubles */
famg_status famg_comm_allgatherv_f64(famg_comm *c, const double *const *local, const int64_t *counts, double *const *out);
/* Row slabs with GLOBAL column ids -> distributed matrix (row offsets follow from the slab heights, rank order).
 * col_split: nranks + 1 column ownership boundaries, NULL for a square operator (same as the row split).
 * The slabs are taken over: famg_dmat_finalize renumbers their columns in place. */
famg_status famg_dmat_create(famg_comm *c, famg_csr *const *slabs, int64_t ncols_global, const int64_t *col_split, famg_dmat **out);
/* collective: builds the halo plan (ghost columns, per-peer send lists) and renumbers the columns of every slab to
 * [owned | ghost] without reordering the entries of a row; replicated_cols != 0 declares the consumer vector
 * replicated instead (global column ids kept, no halo) */
famg_status famg_dmat_finalize(famg_dmat *m, int replicated_cols);
famg_status famg_dmat_retain(famg_dmat *m);
famg_status famg_dmat_destroy(famg_dmat *m);
/* global shape and the nranks + 1 row / column ownership boundaries (any pointer may be NULL) */
famg_status famg_dmat_info(const famg_dmat *m, int64_t *nrows, int64_t *ncols, int64_t *row_split, int64_t *col_split);
/* slab of hosted rank `local_index`; global_cols != 0: a copy with global column ids (verification, download) */
famg_status famg_dmat_local(const famg_dmat *m, int local_index, int global_cols, famg_csr **out);
/* collective: the whole matrix replicated on every rank (out: one handle per hosted rank) -- the transition to
 * the replicated coarse tail */
famg_status famg_dmat_gather(const famg_dmat *m, famg_csr **out);
/* One coarsening step (smoothed_aggregation, interpolation/mod.rs:730-836; block_size 1, one near-null vector):
 * a: finalized square operator.  Per hosted rank: its own aggregates over LOCAL row ids (n_aggs, agg_ptr,
 * agg_nodes -- aggregates never straddle ranks; coarse ids are assigned rank after rank), the local slice of the
 * near-null vector, and coarse_nn (out, n_aggs doubles).  Returns P (rows of a, global coarse column ids, not
 * finalized), R (finalized over a's split) and A_c (global column ids, not finalized). */
famg_status famg_dist_coarsen(famg_dmat *a, const int64_t *n_aggs, const uint64_t *const *agg_ptr, const uint64_t *const *agg_nodes,
                              const double *const *near_null, int smoothing_steps, double omega, famg_dmat **p, famg_dmat **r,
                              famg_dmat **a_coarse, double *const *coarse_nn);
/* the same with device aggregates and device near-null columns (one famg_partition / famg_vec per hosted rank) */
famg_status famg_dist_coarsen_dev(famg_dmat *a, famg_partition *const *parts, const famg_vec *const *near_null, int smoothing_steps,
                                  double omega, famg_dmat **p, famg_dmat **r, famg_dmat **a_coarse, famg_vec *const *coarse_nn);
famg_status famg_dist_smooth_near_null_dev(famg_dmat *a, int iters, famg_vec *const *near_null);
/* hierarchy.rs:217-228 on a finalized distributed level: `iters`-step L1 stationary iteration on the near-null
 * slices (in place, host), then the thin Q of the single column with the sum of squares chained through the ranks
 * in order -- bit-identical to the undistributed build */
famg_status famg_dist_smooth_near_null(famg_dmat *a, int iters, double *const *near_null);
/* Distributed Multigrid + PCG from level-wise distributed operators: a[l], r[l], p[l] (l < nlevels; all finalized,
 * p[nlevels-1] with replicated columns) and the replicated Multigrid `tail` of the remaining levels (borrowed).
 * diag_kind: FAMG_DIAG_L1 | FAMG_DIAG_JACOBI smoother of the distributed levels. */
famg_status famg_dist_mg_create_levels(famg_comm *c, int nlevels, famg_dmat *const *a, famg_dmat *const *r, famg_dmat *const *p,
                                       int diag_kind, double omega, famg_mg *tail, famg_dist_mg **out);

/* ---- instrumentation ---------------------------------------------------------------------- */
/* tuning knobs for A/B measurements: "spmv_variant" (1 = one staged chunk per CTA, 2 = persistent
 * TMA-fed pipeline), "tma_min_rows" (smallest operator the persistent kernel is used for), "spmm_cb"
 * (right-hand sides per row walk for k > 1: 1 | 2), "block_build_host" (1 = famg_smoother_block inverts its blocks on the
 * host).  Changing an option invalidates captured cycle graphs.
 * One solve at a time per context: PCG work vectors and reduction scratch are context-owned. */
famg_status famg_ctx_set_option(famg_ctx *ctx, const char *key, int64_t value);
/* In-kernel timeline: famg_ctx_set_option(ctx, "trace", 1) makes the SpMV-family and halo-exchange kernels launched (or
 * captured) afterwards stamp %globaltimer at their begin / end and at the stages of an exchange; famg_ctx_trace_dump
 * writes the records as text (see csrc/core.cu) and "trace" = 1 again restarts the list.  Diagnostics only. */
famg_status famg_ctx_trace_dump(famg_ctx *ctx, const char *path);
/* Pre-size the context's device memory pool (operators, temporaries and multivectors are carved out of it): one
 * allocation of `bytes` handed straight back to the pool, which keeps the physical memory.  Avoids mapping new physical
 * memory in the middle of a hierarchy build. */
famg_status famg_ctx_reserve(famg_ctx *ctx, int64_t bytes);
/* OpenMP threads of the host-side setup pieces (tentative prolongator, CSR validation, partitioner).  Launchers such as
 * torchrun export OMP_NUM_THREADS=1 to every rank; a rank may claim its share of the host cores here. */
famg_status famg_set_num_threads(int n);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
famg_status famg_ctx_launch_count(const famg_ctx *ctx, int64_t *count);
/* time `reps` back-to-back launches of one fused kernel class with CUDA events on the context
 * stream; returns average milliseconds per launch.  which: 0 SpMV y=Ax, 1 residual, 2 fused diagonal
 * smoother sweep (square operators), 3 y += A x (prolongation-correct).  Rectangular operators are accepted
 * for 0, 1 and 3.  Vectors are internal scratch (values irrelevant). */
famg_status famg_time_kernel(const famg_csr *a, int which, int reps, int warmup, float *ms_avg);

#ifdef __cplusplus
}
#endif
#endif /* FAMG_H */
