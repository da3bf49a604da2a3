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
    // PCG
    pub fn famg_pcg_solve(a: *const famg_csr, pc_kind: c_int, precond: *mut c_void, x: *mut f64, b: *const f64,
                          rel_tol: f64, abs_tol: f64, max_iters: i64, zero_guess: c_int, info: *mut famg_cg_info) -> famg_status;
}
