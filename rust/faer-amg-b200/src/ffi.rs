//! `extern "C"` declarations of include/famg.h (subset used by the wrappers in lib.rs).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub type famg_status = c_int;
#[repr(C)] pub struct famg_ctx { _p: [u8; 0] }
#[repr(C)] pub struct famg_csr { _p: [u8; 0] }
#[repr(C)] pub struct famg_vec { _p: [u8; 0] }
#[repr(C)] pub struct famg_smoother { _p: [u8; 0] }
#[repr(C)] pub struct famg_mg { _p: [u8; 0] }
#[repr(C)] pub struct famg_composite { _p: [u8; 0] }
#[repr(C)] pub struct famg_comm { _p: [u8; 0] }
#[repr(C)] pub struct famg_dmat { _p: [u8; 0] }
#[repr(C)] pub struct famg_dist_mg { _p: [u8; 0] }
#[repr(C)] pub struct famg_partition { _p: [u8; 0] }
pub const FAMG_UNIQUE_ID_BYTES: usize = 128;

#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct famg_cg_info { pub iter_count: i64, pub abs_residual: f64, pub rel_residual: f64 }

pub const FAMG_OK: famg_status = 0;
pub const FAMG_ERR_NO_CONVERGENCE: famg_status = 5;
pub const FAMG_ERR_NOT_SPD: famg_status = 6;
pub const FAMG_PC_NONE: c_int = 0;
pub const FAMG_PC_SMOOTHER: c_int = 1;
pub const FAMG_PC_MG: c_int = 2;
pub const FAMG_PC_COMPOSITE: c_int = 3;

extern "C" {
    pub fn famg_last_error() -> *const c_char;
    pub fn famg_ctx_create(device: c_int, out: *mut *mut famg_ctx) -> famg_status;
    pub fn famg_ctx_destroy(ctx: *mut famg_ctx) -> famg_status;
    // SparseMatOp::new / ParSpmmOp::new
    pub fn famg_csr_create(ctx: *mut famg_ctx, nrows: i64, ncols: i64, row_ptr: *const usize,
                           col_idx: *const usize, val: *const f64, out: *mut *mut famg_csr) -> famg_status;
    pub fn famg_csr_destroy(a: *mut famg_csr) -> famg_status;
    pub fn famg_csr_dims(a: *const famg_csr, nrows: *mut i64, ncols: *mut i64, nnz: *mut i64) -> famg_status;
    pub fn famg_csr_download(a: *const famg_csr, row_ptr: *mut usize, col_idx: *mut usize, val: *mut f64) -> famg_status;
    // LinOp::apply
    pub fn famg_spmm(a: *const famg_csr, out: *mut f64, ld_out: i64, rhs: *const f64, ld_rhs: i64, k: i64) -> famg_status;
    // smoothers
    pub fn famg_smoother_diag(a: *const famg_csr, kind: c_int, omega: f64, out: *mut *mut famg_smoother) -> famg_status;
    pub fn famg_smoother_cholesky(a: *const famg_csr, out: *mut *mut famg_smoother) -> famg_status;
    pub fn famg_smoother_block(a: *const famg_csr, n_aggs: i64, agg_ptr: *const usize, agg_nodes: *const usize,
                               out: *mut *mut famg_smoother) -> famg_status;
    pub fn famg_smoother_apply(s: *const famg_smoother, out: *mut f64, ld_out: i64, rhs: *const f64, ld_rhs: i64, k: i64) -> famg_status;
    pub fn famg_smoother_destroy(s: *mut famg_smoother) -> famg_status;
    // Multigrid
    pub fn famg_mg_create(a0: *const famg_csr, s0: *const famg_smoother, out: *mut *mut famg_mg) -> famg_status;
    pub fn famg_mg_add_level(mg: *mut famg_mg, a: *const famg_csr, s: *const famg_smoother,
                             r: *const famg_csr, p: *const famg_csr) -> famg_status;
    pub fn famg_mg_set_cycle(mg: *mut famg_mg, mu: c_int, smoothing_steps: c_int) -> famg_status;
    pub fn famg_mg_apply(mg: *mut famg_mg, out: *mut f64, ld_out: i64, rhs: *const f64, ld_rhs: i64, k: i64) -> famg_status;
    pub fn famg_mg_destroy(mg: *mut famg_mg) -> famg_status;
    // smoothed_aggregation's sparse expressions
    pub fn famg_galerkin(a: *const famg_csr, p0: *const famg_csr, smoothing_steps: c_int, omega: f64,
                         p: *mut *mut famg_csr, r: *mut *mut famg_csr, a_coarse: *mut *mut famg_csr) -> famg_status;
    pub fn famg_galerkin_block(a: *const famg_csr, p0: *const famg_csr, block_size: i64, smoothing_steps: c_int, omega: f64,
                               p: *mut *mut famg_csr, r: *mut *mut famg_csr, a_coarse: *mut *mut famg_csr) -> famg_status;
    pub fn famg_block_jacobi(a: *const famg_csr, block_size: i64, p: *const famg_csr, out: *mut *mut famg_csr) -> famg_status;
    pub fn famg_smooth_p(a: *const famg_csr, m_inv: *const famg_csr, p: *const famg_csr, out: *mut *mut famg_csr) -> famg_status;
    // block smoother (block_smoothers.rs:88-123); vdim > 1 = diagonally_compensate_vector
    pub fn famg_smoother_block_vector(a: *const famg_csr, vdim: i64, n_aggs: i64, agg_ptr: *const usize,
                                      agg_nodes: *const usize, out: *mut *mut famg_smoother) -> famg_status;
    // near-null search kept on the device (adaptivity.rs:191-198, 307-390)
    pub fn famg_error_propagator_dev(a: *const famg_csr, s: *const famg_smoother, out: *mut famg_vec, x: *const famg_vec) -> famg_status;
    pub fn famg_smooth_vector_pc_dev(a: *const famg_csr, pc_kind: c_int, precond: *mut c_void, iterations: i64,
                                     x: *mut famg_vec, cfs: *mut f64) -> famg_status;
    pub fn famg_thin_q_dev(x: *mut famg_vec) -> famg_status;
    // Composite (preconditioners/composite.rs)
    pub fn famg_composite_create(a: *const famg_csr, out: *mut *mut famg_composite) -> famg_status;
    pub fn famg_composite_push(c: *mut famg_composite, pc_kind: c_int, component: *mut c_void) -> famg_status;
    pub fn famg_composite_apply_dev(c: *mut famg_composite, out: *mut famg_vec, rhs: *const famg_vec) -> famg_status;
    pub fn famg_composite_destroy(c: *mut famg_composite) -> famg_status;
    // the remaining sparse expressions of smoothed_aggregation (interpolation/mod.rs:747-828)
    pub fn famg_spgemm(a: *const famg_csr, b: *const famg_csr, out: *mut *mut famg_csr) -> famg_status;
    pub fn famg_transpose(a: *const famg_csr, out: *mut *mut famg_csr) -> famg_status;
    pub fn famg_smooth_interpolation(a: *const famg_csr, p: *const famg_csr, omega: f64, out: *mut *mut famg_csr) -> famg_status;
    pub fn famg_tentative_p(ctx: *mut famg_ctx, n_fine: i64, block_size: i64, k: i64, cand: i64, near_null: *const f64, ld_nn: i64,
                            n_aggs: i64, agg_ptr: *const usize, agg_nodes: *const usize, p: *mut *mut famg_csr,
                            coarse_nn: *mut f64) -> famg_status;
    pub fn famg_thin_q(n: i64, k: i64, a: *mut f64, lda: i64) -> famg_status;
    // device-resident aggregates / tentative prolongator
    pub fn famg_partition_upload(ctx: *mut famg_ctx, n_nodes: i64, n_aggs: i64, agg_ptr: *const usize, agg_nodes: *const usize,
                                 out: *mut *mut famg_partition) -> famg_status;
    pub fn famg_partition_destroy(p: *mut famg_partition) -> famg_status;
    pub fn famg_tentative_p_dev(part: *const famg_partition, near_null: *const famg_vec, p: *mut *mut famg_csr,
                                coarse_nn: *mut famg_vec) -> famg_status;
    // device vectors
    pub fn famg_vec_create(ctx: *mut famg_ctx, nrows: i64, ncols: i64, out: *mut *mut famg_vec) -> famg_status;
    pub fn famg_vec_destroy(v: *mut famg_vec) -> famg_status;
    pub fn famg_vec_upload(v: *mut famg_vec, host: *const f64, ld: i64) -> famg_status;
    pub fn famg_vec_download(v: *const famg_vec, host: *mut f64, ld: i64) -> famg_status;
    // multi-GPU: one rank per GPU, row slabs (include/famg.h "distributed hierarchy construction")
    pub fn famg_comm_unique_id(id_bytes: *mut c_void) -> famg_status;
    pub fn famg_comm_create(ctx: *mut famg_ctx, nranks: c_int, rank: c_int, id_bytes: *const c_void, out: *mut *mut famg_comm) -> famg_status;
    pub fn famg_comm_destroy(c: *mut famg_comm) -> famg_status;
    pub fn famg_comm_allgatherv_f64(c: *mut famg_comm, local: *const *const f64, counts: *const i64, out: *const *mut f64) -> famg_status;
    pub fn famg_dmat_create(c: *mut famg_comm, slabs: *const *mut famg_csr, ncols_global: i64, col_split: *const i64,
                            out: *mut *mut famg_dmat) -> famg_status;
    pub fn famg_dmat_finalize(m: *mut famg_dmat, replicated_cols: c_int) -> famg_status;
    pub fn famg_dmat_destroy(m: *mut famg_dmat) -> famg_status;
    pub fn famg_dmat_info(m: *const famg_dmat, nrows: *mut i64, ncols: *mut i64, row_split: *mut i64, col_split: *mut i64) -> famg_status;
    pub fn famg_dmat_local(m: *const famg_dmat, local_index: c_int, global_cols: c_int, out: *mut *mut famg_csr) -> famg_status;
    pub fn famg_dmat_gather(m: *const famg_dmat, out: *mut *mut famg_csr) -> famg_status;
    pub fn famg_dist_coarsen(a: *mut famg_dmat, n_aggs: *const i64, agg_ptr: *const *const usize, agg_nodes: *const *const usize,
                             near_null: *const *const f64, smoothing_steps: c_int, omega: f64, p: *mut *mut famg_dmat,
                             r: *mut *mut famg_dmat, a_coarse: *mut *mut famg_dmat, coarse_nn: *const *mut f64) -> famg_status;
    pub fn famg_dist_smooth_near_null(a: *mut famg_dmat, iters: c_int, near_null: *const *mut f64) -> famg_status;
    pub fn famg_dist_mg_create_levels(c: *mut famg_comm, nlevels: c_int, a: *const *mut famg_dmat, r: *const *mut famg_dmat,
                                      p: *const *mut famg_dmat, diag_kind: c_int, omega: f64, tail: *mut famg_mg,
                                      out: *mut *mut famg_dist_mg) -> famg_status;
    pub fn famg_dist_mg_destroy(d: *mut famg_dist_mg) -> famg_status;
    pub fn famg_dist_pcg_solve(d: *mut famg_dist_mg, x_local: *mut f64, b_local: *const f64, rel_tol: f64, abs_tol: f64,
                               max_iters: i64, zero_guess: c_int, info: *mut famg_cg_info) -> famg_status;
    // PCG
    pub fn famg_pcg_solve(a: *const famg_csr, pc_kind: c_int, precond: *mut c_void, x: *mut f64, b: *const f64,
                          rel_tol: f64, abs_tol: f64, max_iters: i64, zero_guess: c_int, info: *mut famg_cg_info) -> famg_status;
}
