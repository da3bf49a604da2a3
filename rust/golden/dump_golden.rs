//! Golden-vector dumper for the REAL crate -- the missing piece that pins this repository's oracle.
//!
//! The reference ships no tests or fixtures and cannot be built in this repository's image (no rustc; faer is an
//! out-of-tree path dependency), so `oracle/` is checked only against itself and analytic identities ("parity
//! unpinned").  A maintainer with the crate's toolchain closes that gap in two minutes:
//!
//!   1. append to `src/interpolation/mod.rs` (the function under test, `smoothed_aggregation`, is private):
//!          #[cfg(test)] mod dump_golden { include!("<path to>/rust/golden/dump_golden.rs"); }
//!   2. FAMG_GOLDEN_DIR=<this repo>/tests/golden cargo test --release dump_golden -- --nocapture
//!   3. in this repository: python -m pytest tests/test_reference_fixtures.py
//!      (compares oracle/ with every tests/golden/ref_*.json bit for bit where the operation order is fixed, and to
//!      1e-12 otherwise; the tests are skipped while no fixture exists).
//!
//! One fixture per case: the operator, the aggregates, the near-null vector, and what the crate computes from them --
//! `smoothed_aggregation` (P, R, A_c, coarse near-null), `new_l1`/`new_jacobi`, a `ParSpmmOp` apply, one `Multigrid`
//! V-cycle with the two-level hierarchy, and the PCG iteration count of `conjugate_gradient` at 1e-8.
//! serde_json is already a dev-dependency of the crate (Cargo.toml:24).
use super::*;
use crate::core::SparseMatOp;
use crate::par_spmm::ParSpmmOp;
use crate::partitioners::Partition;
use crate::preconditioners::multigrid::Multigrid;
use crate::preconditioners::smoothers::{new_jacobi, new_l1};
use faer::sparse::{SparseRowMat, Triplet};
use faer::{Mat, Par};
use serde_json::{json, Value};
use std::collections::BTreeSet;
use std::sync::Arc;

/// G7: 7-point Laplacian, lexicographic i = x + nx (y + ny z), diagonal 6, off-diagonals -1
fn gen_g7(nx: usize, ny: usize, nz: usize) -> SparseRowMat<usize, f64> {
    let n = nx * ny * nz;
    let mut t = Vec::new();
    for z in 0..nz { for y in 0..ny { for x in 0..nx {
        let i = x + nx * (y + ny * z);
        if z > 0 { t.push(Triplet::new(i, i - nx * ny, -1.0)); }
        if y > 0 { t.push(Triplet::new(i, i - nx, -1.0)); }
        if x > 0 { t.push(Triplet::new(i, i - 1, -1.0)); }
        t.push(Triplet::new(i, i, 6.0));
        if x + 1 < nx { t.push(Triplet::new(i, i + 1, -1.0)); }
        if y + 1 < ny { t.push(Triplet::new(i, i + nx, -1.0)); }
        if z + 1 < nz { t.push(Triplet::new(i, i + nx * ny, -1.0)); }
    }}}
    SparseRowMat::try_new_from_triplets(n, n, &t).unwrap()
}

/// G27: Kx(x)My(x)Mz + ey Mx(x)Ky(x)Mz + ez Mx(x)My(x)Kz, K = tridiag(-1,2,-1), M = tridiag(1,4,1)/6; the three terms are
/// formed and added left to right (the generators of this repository round the same way)
fn gen_g27(nx: usize, ny: usize, nz: usize, ey: f64, ez: f64) -> SparseRowMat<usize, f64> {
    let k = |d: i64| if d == 0 { 2.0 } else { -1.0 };
    let m = |d: i64| if d == 0 { 4.0 / 6.0 } else { 1.0 / 6.0 };
    let n = nx * ny * nz;
    let mut t = Vec::new();
    for z in 0..nz as i64 { for y in 0..ny as i64 { for x in 0..nx as i64 {
        let i = (x + nx as i64 * (y + ny as i64 * z)) as usize;
        for dz in -1..=1i64 { if z + dz < 0 || z + dz >= nz as i64 { continue; }
        for dy in -1..=1i64 { if y + dy < 0 || y + dy >= ny as i64 { continue; }
        for dx in -1..=1i64 { if x + dx < 0 || x + dx >= nx as i64 { continue; }
            let t0 = k(dx) * m(dy) * m(dz);
            let t1 = ey * m(dx) * k(dy) * m(dz);
            let t2 = ez * m(dx) * m(dy) * k(dz);
            let j = (i as i64 + dx + nx as i64 * (dy + ny as i64 * dz)) as usize;
            t.push(Triplet::new(i, j, t0 + t1 + t2));
        }}}
    }}}
    SparseRowMat::try_new_from_triplets(n, n, &t).unwrap()
}

/// 2x2x2 boxes, a trailing partial box joins its predecessor; nodes ascending (BTreeSet)
fn geometric_aggregates(nx: usize, ny: usize, nz: usize) -> Vec<BTreeSet<usize>> {
    let (cx, cy, cz) = ((nx / 2).max(1), (ny / 2).max(1), (nz / 2).max(1));
    let mut aggs = vec![BTreeSet::new(); cx * cy * cz];
    for z in 0..nz { for y in 0..ny { for x in 0..nx {
        let (ax, ay, az) = ((x / 2).min(cx - 1), (y / 2).min(cy - 1), (z / 2).min(cz - 1));
        aggs[ax + cx * (ay + cy * az)].insert(x + nx * (y + ny * z));
    }}}
    aggs
}

fn csr_json(m: &SparseRowMat<usize, f64>) -> Value {
    let s = m.symbolic();
    json!({"nrows": m.nrows(), "ncols": m.ncols(), "row_ptr": s.row_ptr(), "col_idx": s.col_idx(),
           // f64 as bit patterns: exact round trip through JSON
           "val_bits": m.val().iter().map(|v| v.to_bits()).collect::<Vec<u64>>()})
}
fn vec_json(v: &[f64]) -> Value { json!(v.iter().map(|x| x.to_bits()).collect::<Vec<u64>>()) }
fn col0(m: &Mat<f64>) -> Vec<f64> { (0..m.nrows()).map(|i| m[(i, 0)]).collect() }

fn dump_case(name: &str, a: SparseRowMat<usize, f64>, dims: (usize, usize, usize)) {
    let n = a.nrows();
    let aggs = geometric_aggregates(dims.0, dims.1, dims.2);
    let partition = Arc::new(Partition::from_agg_to_node(aggs.clone()));
    let near_null = Mat::<f64>::from_fn(n, 1, |_, _| 1.0 / (n as f64).sqrt());
    let (coarse_nn, r, p, ac, _) = smoothed_aggregation(a.as_ref(), partition, 1, near_null.as_ref(), 1, 1);
    // diagonal smoothers
    let l1 = new_l1(&a.as_ref());
    let jac = new_jacobi(&a.as_ref(), 0.66);
    let d_l1: Vec<f64> = (0..n).map(|i| l1.column_vector()[i]).collect();
    let d_jac: Vec<f64> = (0..n).map(|i| jac.column_vector()[i]).collect();
    // ParSpmmOp apply on a reproducible vector (no RNG: x_i = sin(0.37 i) + 1e-3 cos(0.011 i))
    let x = Mat::<f64>::from_fn(n, 1, |i, _| (0.37 * i as f64).sin() + 1e-3 * (0.011 * i as f64).cos());
    let arc_a = Arc::new(a);
    let op = SparseMatOp::new(arc_a.clone(), 1);
    let mut y = Mat::<f64>::zeros(n, 1);
    let mut buf = faer::dyn_stack::MemBuffer::new(faer::dyn_stack::StackReq::empty());
    let stack = faer::dyn_stack::MemStack::new(&mut buf);
    op.dyn_op().apply(y.as_mut(), x.as_ref(), Par::Seq, stack);
    // two-level V(1,1) cycle, L1 smoother, exact coarse solve; PCG to 1e-8 with b = 1
    let (arc_r, arc_p, arc_ac) = (Arc::new(r), Arc::new(p), Arc::new(ac));
    let mut mg = Multigrid::new(arc_a.clone(), Arc::new(new_l1(&arc_a.as_ref().as_ref())));
    let coarse = crate::preconditioners::coarse_solvers::SparseCholeskySolve::new(arc_ac.as_ref().as_ref());
    mg.add_level(arc_ac.clone(), Arc::new(coarse), arc_r.clone(), arc_p.clone());
    let mut z = Mat::<f64>::zeros(n, 1);
    mg.apply(z.as_mut(), x.as_ref(), Par::Seq, stack);
    let b = Mat::<f64>::from_fn(n, 1, |_, _| 1.0);
    let ((iters, rel_res, _), _) = crate::utils::test_solver(op.dyn_op(), Arc::new(mg), None, Some(b.as_ref()), 1000, 1e-8);
    let out = json!({
        "case": name, "dims": [dims.0, dims.1, dims.2], "a": csr_json(&arc_a),
        "agg_ptr": aggs.iter().scan(0usize, |s, g| { *s += g.len(); Some(*s) }).collect::<Vec<_>>(),
        "agg_nodes": aggs.iter().flat_map(|g| g.iter().copied()).collect::<Vec<_>>(),
        "p": csr_json(&arc_p), "r": csr_json(&arc_r), "ac": csr_json(&arc_ac), "coarse_nn": vec_json(&col0(&coarse_nn)),
        "l1": vec_json(&d_l1), "jacobi_0.66": vec_json(&d_jac), "x": vec_json(&col0(&x)), "a_times_x": vec_json(&col0(&y)),
        "two_level_vcycle_of_x": vec_json(&col0(&z)), "pcg": {"rel_tol": 1e-8, "iters": iters, "rel_residual": rel_res},
    });
    let dir = std::env::var("FAMG_GOLDEN_DIR").expect("set FAMG_GOLDEN_DIR to <faer-amg-b200>/tests/golden");
    std::fs::write(format!("{dir}/ref_{name}.json"), serde_json::to_string(&out).unwrap()).unwrap();
    println!("wrote {dir}/ref_{name}.json ({n} rows, pcg {iters} iterations)");
}

#[test]
fn dump_golden() {
    // ParSpmmOp refuses Par::Seq (par_spmm.rs:35); two workers are enough for the tiled path, whose result does not depend on
    // the worker count (every output entry accumulates in ascending column order)
    faer::set_global_parallelism(Par::rayon(2));
    dump_case("g7_8", gen_g7(8, 8, 8), (8, 8, 8));
    dump_case("g7_12x10x7", gen_g7(12, 10, 7), (12, 10, 7));
    dump_case("g27_8", gen_g27(8, 8, 8, 1.0, 1e-2), (8, 8, 8));
    dump_case("g7_24", gen_g7(24, 24, 24), (24, 24, 24));
}
