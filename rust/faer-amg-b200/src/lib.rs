//! Drop-in GPU operators for faer-amg: the types below implement faer's `LinOp` / `Precond`
//! traits by calling libfamg.so, so `SparseMatOp::dyn_op()` (src/core.rs:88-92) and
//! `Multigrid::add_level` (src/preconditioners/multigrid.rs:228-239) can hand them to every
//! existing consumer unchanged.  Errors keep the crate's convention: a non-zero status panics
//! with `famg_last_error()` (cf. par_spmm.rs:35,78; hierarchy.rs:259-264).
pub mod ffi;

use faer::dyn_stack::{MemStack, StackReq};
use faer::matrix_free::{BiLinOp, BiPrecond, LinOp, Precond};
use faer::sparse::SparseRowMatRef;
use faer::{MatMut, MatRef, Par};
use std::ffi::CStr;
use std::sync::Arc;

fn check(st: ffi::famg_status) {
    if st != ffi::FAMG_OK {
        let msg = unsafe { CStr::from_ptr(ffi::famg_last_error()) }.to_string_lossy().into_owned();
        panic!("libfamg: {msg}");
    }
}

/// One per (process, device); replaces `faer::set_global_parallelism(Par::Rayon(n))`.
#[derive(Debug)]
pub struct GpuContext(pub(crate) *mut ffi::famg_ctx);
unsafe impl Send for GpuContext {}
unsafe impl Sync for GpuContext {}
impl GpuContext {
    pub fn new(device: i32) -> Arc<Self> {
        let mut p = std::ptr::null_mut();
        check(unsafe { ffi::famg_ctx_create(device, &mut p) });
        Arc::new(Self(p))
    }
}
impl Drop for GpuContext { fn drop(&mut self) { unsafe { ffi::famg_ctx_destroy(self.0); } } }

/// GPU replacement of `ParSpmmOp` (src/par_spmm.rs:17-159).
#[derive(Debug)]
pub struct GpuSpmmOp { h: *mut ffi::famg_csr, nrows: usize, ncols: usize, _ctx: Arc<GpuContext> }
unsafe impl Send for GpuSpmmOp {}
unsafe impl Sync for GpuSpmmOp {}

impl GpuSpmmOp {
    /// Same signature role as `ParSpmmOp::new(mat, par)`: copies the (compressed) CSR to HBM.
    pub fn new(ctx: Arc<GpuContext>, mat: SparseRowMatRef<usize, f64>) -> Self {
        let sym = mat.symbolic();
        assert!(sym.row_nnz().is_none(), "matrix must be in compressed mode");
        let mut h = std::ptr::null_mut();
        check(unsafe {
            ffi::famg_csr_create(ctx.0, mat.nrows() as i64, mat.ncols() as i64, sym.row_ptr().as_ptr(),
                                 sym.col_idx().as_ptr(), mat.val().as_ptr(), &mut h)
        });
        Self { h, nrows: mat.nrows(), ncols: mat.ncols(), _ctx: ctx }
    }
    pub(crate) fn raw(&self) -> *const ffi::famg_csr { self.h }
}
impl Drop for GpuSpmmOp { fn drop(&mut self) { unsafe { ffi::famg_csr_destroy(self.h); } } }

impl LinOp<f64> for GpuSpmmOp {
    fn apply_scratch(&self, _rhs_ncols: usize, _par: Par) -> StackReq { StackReq::empty() } // par_spmm.rs:144-149
    fn nrows(&self) -> usize { self.nrows }
    fn ncols(&self) -> usize { self.ncols }
    fn apply(&self, out: MatMut<'_, f64>, rhs: MatRef<'_, f64>, _par: Par, _stack: &mut MemStack) {
        assert_eq!(out.row_stride(), 1); assert_eq!(rhs.row_stride(), 1);
        let mut out = out;
        check(unsafe {
            ffi::famg_spmm(self.h, out.as_ptr_mut(), out.col_stride() as i64, rhs.as_ptr(), rhs.col_stride() as i64,
                           rhs.ncols() as i64)
        });
    }
    fn conj_apply(&self, out: MatMut<'_, f64>, rhs: MatRef<'_, f64>, par: Par, stack: &mut MemStack) {
        self.apply(out, rhs, par, stack) // par_spmm.rs:156-158
    }
}

/// GPU replacement of the level smoothers (`Diag` from new_l1/new_l2/new_jacobi,
/// `SparseCholeskySolve`, `BlockSmoother`): anything `Multigrid` stores as `Arc<dyn BiPrecond<f64>>`.
#[derive(Debug)]
pub struct GpuSmoother { h: *mut ffi::famg_smoother, n: usize }
unsafe impl Send for GpuSmoother {}
unsafe impl Sync for GpuSmoother {}
impl GpuSmoother {
    pub fn new_l1(op: &GpuSpmmOp) -> Self { Self::diag(op, 0, 0.0) }       // smoothers.rs:63-76
    pub fn new_l2(op: &GpuSpmmOp) -> Self { Self::diag(op, 1, 0.0) }       // smoothers.rs:43-61
    pub fn new_jacobi(op: &GpuSpmmOp, omega: f64) -> Self { Self::diag(op, 2, omega) } // :78-86
    fn diag(op: &GpuSpmmOp, kind: i32, omega: f64) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { ffi::famg_smoother_diag(op.raw(), kind, omega, &mut h) });
        Self { h, n: op.nrows }
    }
    pub fn cholesky(op: &GpuSpmmOp) -> Self {                              // coarse_solvers.rs:173-180
        let mut h = std::ptr::null_mut();
        check(unsafe { ffi::famg_smoother_cholesky(op.raw(), &mut h) });
        Self { h, n: op.nrows }
    }
    /// `BlockSmoother::new` (block_smoothers.rs:88-123): one diagonally compensated block per aggregate,
    /// solved exactly; `vdim > 1` lumps off-aggregate coupling blocks by `diagonally_compensate_vector`.
    /// `agg_ptr` / `agg_nodes`: the `Partition`'s aggregates in CSR form (BTreeSet order).
    pub fn block(op: &GpuSpmmOp, vdim: usize, agg_ptr: &[usize], agg_nodes: &[usize]) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe {
            ffi::famg_smoother_block_vector(op.raw(), vdim as i64, agg_ptr.len() as i64 - 1, agg_ptr.as_ptr(), agg_nodes.as_ptr(), &mut h)
        });
        Self { h, n: op.nrows }
    }
    pub(crate) fn raw(&self) -> *const ffi::famg_smoother { self.h }
}
impl Drop for GpuSmoother { fn drop(&mut self) { unsafe { ffi::famg_smoother_destroy(self.h); } } }
impl LinOp<f64> for GpuSmoother {
    fn apply_scratch(&self, _: usize, _: Par) -> StackReq { StackReq::empty() }
    fn nrows(&self) -> usize { self.n }
    fn ncols(&self) -> usize { self.n }
    fn apply(&self, out: MatMut<'_, f64>, rhs: MatRef<'_, f64>, _: Par, _: &mut MemStack) {
        let mut out = out;
        check(unsafe {
            ffi::famg_smoother_apply(self.h, out.as_ptr_mut(), out.col_stride() as i64, rhs.as_ptr(),
                                     rhs.col_stride() as i64, rhs.ncols() as i64)
        });
    }
    fn conj_apply(&self, out: MatMut<'_, f64>, rhs: MatRef<'_, f64>, par: Par, stack: &mut MemStack) { self.apply(out, rhs, par, stack) }
}
impl BiLinOp<f64> for GpuSmoother {
    fn transpose_apply_scratch(&self, n: usize, par: Par) -> StackReq { self.apply_scratch(n, par) }
    fn transpose_apply(&self, out: MatMut<'_, f64>, rhs: MatRef<'_, f64>, par: Par, stack: &mut MemStack) { self.apply(out, rhs, par, stack) }
    fn adjoint_apply(&self, out: MatMut<'_, f64>, rhs: MatRef<'_, f64>, par: Par, stack: &mut MemStack) { self.apply(out, rhs, par, stack) }
}
impl Precond<f64> for GpuSmoother {}
impl BiPrecond<f64> for GpuSmoother {}

/// GPU `Multigrid` (src/preconditioners/multigrid.rs:171-249): same builder surface.
#[derive(Debug)]
pub struct GpuMultigrid { h: *mut ffi::famg_mg, n: usize }
unsafe impl Send for GpuMultigrid {}
unsafe impl Sync for GpuMultigrid {}
impl GpuMultigrid {
    pub fn new(op: &GpuSpmmOp, smoother: &GpuSmoother) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { ffi::famg_mg_create(op.raw(), smoother.raw(), &mut h) });
        Self { h, n: op.nrows }
    }
    pub fn add_level(&mut self, op: &GpuSpmmOp, smoother: &GpuSmoother, r: &GpuSpmmOp, p: &GpuSpmmOp) {
        check(unsafe { ffi::famg_mg_add_level(self.h, op.raw(), smoother.raw(), r.raw(), p.raw()) });
    }
    pub fn with_cycle(self, mu: usize, smoothing_steps: usize) -> Self {
        assert!(mu > 0 && smoothing_steps > 0);
        check(unsafe { ffi::famg_mg_set_cycle(self.h, mu as i32, smoothing_steps as i32) });
        self
    }
    pub(crate) fn raw(&self) -> *mut ffi::famg_mg { self.h }
}
impl Drop for GpuMultigrid { fn drop(&mut self) { unsafe { ffi::famg_mg_destroy(self.h); } } }
impl LinOp<f64> for GpuMultigrid {
    fn apply_scratch(&self, _: usize, _: Par) -> StackReq { StackReq::empty() }
    fn nrows(&self) -> usize { self.n }
    fn ncols(&self) -> usize { self.n }
    fn apply(&self, out: MatMut<'_, f64>, rhs: MatRef<'_, f64>, _: Par, _: &mut MemStack) {
        let mut out = out;
        check(unsafe {
            ffi::famg_mg_apply(self.h, out.as_ptr_mut(), out.col_stride() as i64, rhs.as_ptr(), rhs.col_stride() as i64,
                               rhs.ncols() as i64)
        });
    }
    fn conj_apply(&self, out: MatMut<'_, f64>, rhs: MatRef<'_, f64>, par: Par, stack: &mut MemStack) { self.apply(out, rhs, par, stack) }
}
impl BiLinOp<f64> for GpuMultigrid {
    fn transpose_apply_scratch(&self, n: usize, par: Par) -> StackReq { self.apply_scratch(n, par) }
    fn transpose_apply(&self, out: MatMut<'_, f64>, rhs: MatRef<'_, f64>, par: Par, stack: &mut MemStack) { self.apply(out, rhs, par, stack) }
    fn adjoint_apply(&self, out: MatMut<'_, f64>, rhs: MatRef<'_, f64>, par: Par, stack: &mut MemStack) { self.apply(out, rhs, par, stack) }
}
impl Precond<f64> for GpuMultigrid {}
impl BiPrecond<f64> for GpuMultigrid {}

/// GPU `Composite` (src/preconditioners/composite.rs:11-100): symmetric multiplicative combination of
/// multigrids around one operator.  Components are kept alive by the `Arc`s held here.
#[derive(Debug)]
pub struct GpuComposite { h: *mut ffi::famg_composite, n: usize, components: Vec<Arc<GpuMultigrid>> }
unsafe impl Send for GpuComposite {}
unsafe impl Sync for GpuComposite {}
impl GpuComposite {
    pub fn new(mat: &GpuSpmmOp, first_component: Arc<GpuMultigrid>) -> Self {           // composite.rs:48-56
        let mut h = std::ptr::null_mut();
        check(unsafe { ffi::famg_composite_create(mat.raw(), &mut h) });
        let mut c = Self { h, n: mat.nrows, components: Vec::new() };
        c.push(first_component);
        c
    }
    pub fn push(&mut self, component: Arc<GpuMultigrid>) {                                // composite.rs:85-87
        check(unsafe { ffi::famg_composite_push(self.h, ffi::FAMG_PC_MG, component.raw() as *mut _) });
        self.components.push(component);
    }
    pub fn components(&self) -> &Vec<Arc<GpuMultigrid>> { &self.components }
    pub(crate) fn raw(&self) -> *mut ffi::famg_composite { self.h }
}
impl Drop for GpuComposite { fn drop(&mut self) { unsafe { ffi::famg_composite_destroy(self.h); } } }

/// Prolongator smoothing for `block_size > 1` (interpolation/mod.rs:963-1028): returns the device CSR of
/// `P - 0.66 D_b^-1 A P`; download with `famg_csr_download` into `SparseRowMat::new(..)`.
pub fn block_jacobi(mat: &GpuSpmmOp, block_size: usize, p: &GpuSpmmOp) -> *mut ffi::famg_csr {
    let mut out = std::ptr::null_mut();
    check(unsafe { ffi::famg_block_jacobi(mat.raw(), block_size as i64, p.raw(), &mut out) });
    out
}

/// Device-resident PCG: the whole `conjugate_gradient(..)` call of utils.rs:600-609 in one FFI
/// call, so vectors cross PCIe once per solve instead of twice per operator apply.
pub fn solve_pcg_device(op: &GpuSpmmOp, pc: &GpuMultigrid, x: &mut [f64], b: &[f64], rel_tol: f64, abs_tol: f64,
                        max_iters: usize, zero_guess: bool) -> Result<ffi::famg_cg_info, ffi::famg_cg_info> {
    let mut info = ffi::famg_cg_info::default();
    let st = unsafe {
        ffi::famg_pcg_solve(op.raw(), ffi::FAMG_PC_MG, pc.raw() as *mut _, x.as_mut_ptr(), b.as_ptr(), rel_tol, abs_tol,
                            max_iters as i64, zero_guess as i32, &mut info)
    };
    match st {
        ffi::FAMG_OK => Ok(info),
        ffi::FAMG_ERR_NO_CONVERGENCE | ffi::FAMG_ERR_NOT_SPD => Err(info), // CgError::NoConvergence{..} etc.
        other => { check(other); unreachable!() }
    }
}
