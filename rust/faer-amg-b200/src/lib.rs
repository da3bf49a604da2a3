//! Drop-in GPU operators for faer-amg: the types below implement faer's `LinOp` / `Precond`
//! traits by calling libfamg.so, so `SparseMatOp::dyn_op()` (src/core.rs:88-92) and
//! `Multigrid::add_level` (src/preconditioners/multigrid.rs:228-239) can hand them to every
//! existing consumer unchanged.  Errors keep the crate's convention: a non-zero status panics
//! with `famg_last_error()` (cf. par_spmm.rs:35,78; hierarchy.rs:259-264).
pub mod ffi;

use faer::dyn_stack::{MemStack, StackReq};
use faer::matrix_free::{BiLinOp, BiPrecond, LinOp, Precond};
use faer::sparse::{SparseRowMat, SparseRowMatRef, SymbolicSparseRowMat};
use faer::{Mat, MatMut, MatRef, Par};
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

// ---------------------------------------------------------------------------------------------------------------
// Plug point P3: the three sparse expressions of `smoothed_aggregation` (interpolation/mod.rs:747-828)
// ---------------------------------------------------------------------------------------------------------------
/// Owned device CSR handle -> faer `SparseRowMat` (the buffers Rust wraps are Rust's own: two-phase download).
pub fn download_csr(h: *mut ffi::famg_csr) -> SparseRowMat<usize, f64> {
    let (mut nr, mut nc, mut nnz) = (0i64, 0i64, 0i64);
    check(unsafe { ffi::famg_csr_dims(h, &mut nr, &mut nc, &mut nnz) });
    let mut row_ptr = vec![0usize; nr as usize + 1];
    let mut col_idx = vec![0usize; nnz as usize];
    let mut val = vec![0f64; nnz as usize];
    check(unsafe { ffi::famg_csr_download(h, row_ptr.as_mut_ptr(), col_idx.as_mut_ptr(), val.as_mut_ptr()) });
    unsafe { ffi::famg_csr_destroy(h) };
    let sym = SymbolicSparseRowMat::new_checked(nr as usize, nc as usize, row_ptr, None, col_idx);
    SparseRowMat::new(sym, val)
}

/// `smoothed_aggregation` (interpolation/mod.rs:730-836) with the tentative prolongator, its smoothing, the transpose
/// and `R * (A * P)` on the GPU.  Same inputs, same outputs `(coarse_near_null, R, P, A_c)`: a drop-in for the body of
/// the reference function (the `PartitionType` is the caller's own `partition`).
pub fn smoothed_aggregation_gpu(
    ctx: &Arc<GpuContext>,
    fine_mat: SparseRowMatRef<usize, f64>,
    aggregates: &[std::collections::BTreeSet<usize>],   // Partition::aggregates()
    block_size: usize,
    near_null: MatRef<f64>,
    candidate_dimension: usize,
    smoothing_steps: usize,
) -> (Mat<f64>, SparseRowMat<usize, f64>, SparseRowMat<usize, f64>, SparseRowMat<usize, f64>) {
    let n_fine = fine_mat.nrows();
    assert_eq!(n_fine % block_size, 0);                                   // :743
    let mut agg_ptr = Vec::with_capacity(aggregates.len() + 1);
    let mut agg_nodes = Vec::with_capacity(n_fine / block_size);
    agg_ptr.push(0usize);
    for agg in aggregates { agg_nodes.extend(agg.iter().copied()); agg_ptr.push(agg_nodes.len()); }   // BTreeSet: ascending
    assert_eq!(near_null.row_stride(), 1);
    let k = near_null.ncols();
    let n_coarse = aggregates.len() * candidate_dimension;
    let mut coarse_nn = Mat::<f64>::zeros(n_coarse, k);
    let a = GpuSpmmOp::new(ctx.clone(), fine_mat);
    let mut p0 = std::ptr::null_mut();
    check(unsafe {
        ffi::famg_tentative_p(ctx.0, n_fine as i64, block_size as i64, k as i64, candidate_dimension as i64, near_null.as_ptr(),
                              near_null.col_stride() as i64, aggregates.len() as i64, agg_ptr.as_ptr(), agg_nodes.as_ptr(), &mut p0,
                              coarse_nn.as_ptr_mut())
    });
    let (mut p, mut r, mut ac) = (std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut());
    check(unsafe { ffi::famg_galerkin_block(a.raw(), p0, block_size as i64, smoothing_steps as i32, 0.66, &mut p, &mut r, &mut ac) });
    unsafe { ffi::famg_csr_destroy(p0) };
    (coarse_nn, download_csr(r), download_csr(p), download_csr(ac))
}

/// `&a * &b` on `SparseRowMat` (interpolation/mod.rs:720,828,938) -- structural, sorted, unpruned; ascending inner index.
pub fn spgemm_gpu(a: &GpuSpmmOp, b: &GpuSpmmOp) -> SparseRowMat<usize, f64> {
    let mut out = std::ptr::null_mut();
    check(unsafe { ffi::famg_spgemm(a.raw(), b.raw(), &mut out) });
    download_csr(out)
}
/// `p.transpose().to_row_major()` (interpolation/mod.rs:824-827)
pub fn transpose_gpu(a: &GpuSpmmOp) -> SparseRowMat<usize, f64> {
    let mut out = std::ptr::null_mut();
    check(unsafe { ffi::famg_transpose(a.raw(), &mut out) });
    download_csr(out)
}

/// `BiLinOp` for the operator itself: the transposed operator is built on the device at construction
/// (`GpuSpmmOp::with_transpose`) -- `StationaryIteration::transpose_apply` (smoothers.rs:179-197) needs it.
#[derive(Debug)]
pub struct GpuBiSpmmOp { pub op: GpuSpmmOp, t: *mut ffi::famg_csr }
unsafe impl Send for GpuBiSpmmOp {}
unsafe impl Sync for GpuBiSpmmOp {}
impl GpuBiSpmmOp {
    pub fn new(op: GpuSpmmOp) -> Self {
        let mut t = std::ptr::null_mut();
        check(unsafe { ffi::famg_transpose(op.raw(), &mut t) });
        Self { op, t }
    }
}
impl Drop for GpuBiSpmmOp { fn drop(&mut self) { unsafe { ffi::famg_csr_destroy(self.t); } } }
impl LinOp<f64> for GpuBiSpmmOp {
    fn apply_scratch(&self, n: usize, par: Par) -> StackReq { self.op.apply_scratch(n, par) }
    fn nrows(&self) -> usize { self.op.nrows() }
    fn ncols(&self) -> usize { self.op.ncols() }
    fn apply(&self, out: MatMut<f64>, rhs: MatRef<f64>, par: Par, stack: &mut MemStack) { self.op.apply(out, rhs, par, stack) }
    fn conj_apply(&self, out: MatMut<f64>, rhs: MatRef<f64>, par: Par, stack: &mut MemStack) { self.op.apply(out, rhs, par, stack) }
}
impl BiLinOp<f64> for GpuBiSpmmOp {
    fn transpose_apply_scratch(&self, _n: usize, _par: Par) -> StackReq { StackReq::empty() }
    fn transpose_apply(&self, mut out: MatMut<f64>, rhs: MatRef<f64>, _par: Par, _stack: &mut MemStack) {
        assert!(out.row_stride() == 1 && rhs.row_stride() == 1);
        check(unsafe { ffi::famg_spmm(self.t, out.as_ptr_mut(), out.col_stride() as i64, rhs.as_ptr(), rhs.col_stride() as i64, rhs.ncols() as i64) });
    }
    fn adjoint_apply(&self, out: MatMut<f64>, rhs: MatRef<f64>, par: Par, stack: &mut MemStack) { self.transpose_apply(out, rhs, par, stack) }
}

// ---------------------------------------------------------------------------------------------------------------
// Multi-GPU: one rank (process or thread) per GPU, row slabs.  `Hierarchy::coarsen` (hierarchy.rs:190-248)
// on slabs: every rank owns the rows [row_split[rank], row_split[rank + 1]) of every level.
// ---------------------------------------------------------------------------------------------------------------
pub struct GpuComm { h: *mut ffi::famg_comm, pub nranks: usize, pub rank: usize, ctx: Arc<GpuContext> }
impl GpuComm {
    /// `id`: 128 bytes from `GpuComm::unique_id()` on rank 0, broadcast by the caller's launcher (MPI, sci-bevy-comm, files ...)
    pub fn new(ctx: Arc<GpuContext>, nranks: usize, rank: usize, id: &[u8; ffi::FAMG_UNIQUE_ID_BYTES]) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { ffi::famg_comm_create(ctx.0, nranks as i32, rank as i32, id.as_ptr() as *const _, &mut h) });
        Self { h, nranks, rank, ctx }
    }
    pub fn unique_id() -> [u8; ffi::FAMG_UNIQUE_ID_BYTES] {
        let mut id = [0u8; ffi::FAMG_UNIQUE_ID_BYTES];
        check(unsafe { ffi::famg_comm_unique_id(id.as_mut_ptr() as *mut _) });
        id
    }
}
impl Drop for GpuComm { fn drop(&mut self) { unsafe { ffi::famg_comm_destroy(self.h); } } }

/// Row-partitioned operator: this rank's slab carries GLOBAL column indices (exactly the rows a rank would cut out of the
/// crate's `SparseRowMat`); `finalize` builds the halo plan.
pub struct GpuDistMat { h: *mut ffi::famg_dmat }
impl GpuDistMat {
    pub fn from_slab(comm: &GpuComm, slab: SparseRowMatRef<usize, f64>, ncols_global: usize) -> Self {
        let op = GpuSpmmOp::new(comm.ctx.clone(), slab);
        let slabs = [op.raw() as *mut ffi::famg_csr];
        let mut h = std::ptr::null_mut();
        check(unsafe { ffi::famg_dmat_create(comm.h, slabs.as_ptr(), ncols_global as i64, std::ptr::null(), &mut h) });
        Self { h }   // the distributed matrix holds its own reference on the slab
    }
    pub fn finalize(self, replicated_cols: bool) -> Self { check(unsafe { ffi::famg_dmat_finalize(self.h, replicated_cols as i32) }); self }
    pub fn nrows_global(&self) -> usize { let mut n = 0i64; check(unsafe { ffi::famg_dmat_info(self.h, &mut n, std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut()) }); n as usize }
    /// the whole operator on every rank (transition to the replicated coarse tail)
    pub fn gather(&self) -> SparseRowMat<usize, f64> {
        let mut out = [std::ptr::null_mut::<ffi::famg_csr>()];
        check(unsafe { ffi::famg_dmat_gather(self.h, out.as_mut_ptr()) });
        download_csr(out[0])
    }
}
impl Drop for GpuDistMat { fn drop(&mut self) { unsafe { ffi::famg_dmat_destroy(self.h); } } }

/// One level of `Hierarchy::coarsen` on slabs: this rank's own aggregates over LOCAL row ids and its slice of the
/// near-null vector in, `(P, R, A_c, coarse near-null slice)` out; `P` / `A_c` still need `finalize`.
pub fn dist_coarsen(a: &GpuDistMat, aggregates: &[std::collections::BTreeSet<usize>], near_null_local: &[f64], smoothing_steps: usize)
    -> (GpuDistMat, GpuDistMat, GpuDistMat, Vec<f64>) {
    let mut agg_ptr = vec![0usize];
    let mut agg_nodes = Vec::new();
    for agg in aggregates { agg_nodes.extend(agg.iter().copied()); agg_ptr.push(agg_nodes.len()); }
    let n_aggs = [aggregates.len() as i64];
    let mut coarse_nn = vec![0f64; aggregates.len()];
    let (mut p, mut r, mut ac) = (std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut());
    check(unsafe {
        ffi::famg_dist_coarsen(a.h, n_aggs.as_ptr(), [agg_ptr.as_ptr()].as_ptr(), [agg_nodes.as_ptr()].as_ptr(), [near_null_local.as_ptr()].as_ptr(),
                               smoothing_steps as i32, 0.66, &mut p, &mut r, &mut ac, [coarse_nn.as_mut_ptr()].as_ptr())
    });
    (GpuDistMat { h: p }, GpuDistMat { h: r }, GpuDistMat { h: ac }, coarse_nn)
}
/// hierarchy.rs:217-228 on a distributed level (3-step L1 stationary iteration, thin Q of the single column)
pub fn dist_smooth_near_null(a: &GpuDistMat, nn_local: &mut [f64]) {
    check(unsafe { ffi::famg_dist_smooth_near_null(a.h, 3, [nn_local.as_mut_ptr()].as_ptr()) });
}

/// Row-partitioned `Multigrid` + PCG: distributed levels from `dist_coarsen`, the replicated tail as a `GpuMultigrid`.
pub struct GpuDistMultigrid { h: *mut ffi::famg_dist_mg, _tail: GpuMultigrid }
impl GpuDistMultigrid {
    pub fn new(comm: &GpuComm, a: &[&GpuDistMat], r: &[&GpuDistMat], p: &[&GpuDistMat], tail: GpuMultigrid) -> Self {
        let (ah, rh, ph): (Vec<_>, Vec<_>, Vec<_>) = (a.iter().map(|m| m.h).collect(), r.iter().map(|m| m.h).collect(), p.iter().map(|m| m.h).collect());
        let mut h = std::ptr::null_mut();
        check(unsafe { ffi::famg_dist_mg_create_levels(comm.h, a.len() as i32, ah.as_ptr(), rh.as_ptr(), ph.as_ptr(), 0 /* L1 */, 0.66, tail.raw(), &mut h) });
        Self { h, _tail: tail }
    }
    /// `conjugate_gradient` on the partitioned system: `x_local` / `b_local` are this rank's rows
    pub fn solve(&self, x_local: &mut [f64], b_local: &[f64], rel_tol: f64, max_iters: usize) -> Result<ffi::famg_cg_info, ffi::famg_cg_info> {
        let mut info = ffi::famg_cg_info::default();
        let st = unsafe { ffi::famg_dist_pcg_solve(self.h, x_local.as_mut_ptr(), b_local.as_ptr(), rel_tol, 0.0, max_iters as i64, 1, &mut info) };
        match st { ffi::FAMG_OK => Ok(info), ffi::FAMG_ERR_NO_CONVERGENCE | ffi::FAMG_ERR_NOT_SPD => Err(info), other => { check(other); unreachable!() } }
    }
}
impl Drop for GpuDistMultigrid { fn drop(&mut self) { unsafe { ffi::famg_dist_mg_destroy(self.h); } } }

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
