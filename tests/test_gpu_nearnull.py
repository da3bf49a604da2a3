"""GPU parity tests of the device-resident near-null search (adaptivity.rs:264-390,
434-443) and of hierarchies built from the algebraic host partitioner (8f-3), through the C ABI,
against the CPU oracle.

Tolerances (f64): the fused E apply with a Diag preconditioner keeps the oracle's operation order
and is compared bit for bit.  The device thin Q is CholeskyQR2, the oracle's a Householder QR: both
produce *the* thin Q with R's diagonal positive, equal up to rounding amplified by cond(X)
(<= 1e-12 absolute on entries of unit-norm columns for the well-conditioned blocks used here);
after `iterations` E-apply + QR steps the tolerance is 1e-9."""
import numpy as np
import pytest

import oracle as O
from oracle import partitioner as OP
from util import to_dev, to_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    import faer_amg_b200 as F
    return F


def _thin_q_dev(F, ctx, m):
    from faer_amg_b200.hierarchy import thin_q_dev
    x = F.DeviceMat.from_host(ctx, m)
    thin_q_dev(x)
    return x.to_host()


@pytest.mark.parametrize("n,k", [(3000, 1), (3000, 5), (4099, 17), (70000, 64), (64, 64), (100, 33)])
def test_thin_q_dev_matches_householder(ctx, F, n, k):
    m = np.random.default_rng(n + k).standard_normal((n, k))
    q = _thin_q_dev(F, ctx, m)
    want = O.thin_q(m)
    assert np.max(np.abs(q.T @ q - np.eye(k))) < 1e-13
    assert np.max(np.abs(q - want)) < 1e-12
    r = q.T @ m  # upper triangular with a positive diagonal
    assert np.all(np.diag(r) > 0) and np.max(np.abs(np.tril(r, -1))) < 1e-10


def test_thin_q_dev_errors_and_degenerate(ctx, F):
    from faer_amg_b200.hierarchy import thin_q_dev
    with pytest.raises(F.FamgError) as e:
        thin_q_dev(F.DeviceMat.from_host(ctx, np.ones((100, 65))))
    assert e.value.status == F._ffi.ERR_UNSUPPORTED
    with pytest.raises(F.FamgError) as e:
        thin_q_dev(F.DeviceMat.from_host(ctx, np.ones((100, 2))))  # rank 1
    assert e.value.status == F._ffi.ERR_NUMERIC
    with pytest.raises(F.FamgError):
        thin_q_dev(F.DeviceMat.from_host(ctx, np.ones((3, 5))))
    thin_q_dev(F.DeviceMat(ctx, 0, 0))
    one = _thin_q_dev(F, ctx, np.array([[3.0], [-4.0]]))
    assert np.allclose(one[:, 0], [0.6, -0.8], atol=1e-15)


def test_coldot_and_create_weights(ctx, F):
    from faer_amg_b200._ffi import call
    from faer_amg_b200.core import _f
    rng = np.random.default_rng(2)
    x, y = rng.standard_normal((50001, 7)), rng.standard_normal((50001, 7))
    out = np.zeros(7)
    dx, dy = F.DeviceMat.from_host(ctx, x), F.DeviceMat.from_host(ctx, y)  # keep the handles alive across the call
    call("famg_vec_coldot", dx._h, dy._h, _f(out))
    want = np.einsum("ij,ij->j", x, y)
    assert np.max(np.abs(out - want) / np.einsum("ij,ij->j", np.abs(x), np.abs(y))) < 1e-14
    o = O.gen_g7(9, 8, 7)
    d = to_dev(ctx, o)
    nn = rng.standard_normal((o.nrows, 3))
    got = F.adaptivity.create_weights(nn, d)
    assert np.allclose(got, O.create_weights(o, nn), rtol=1e-13, atol=0)  # adaptivity.rs:434-443


@pytest.mark.parametrize("gen,dims,k", [(O.gen_g7, (10, 9, 8), 4), (O.gen_g27, (7, 6, 5), 3)])
def test_error_propagator_fused_and_block(ctx, F, gen, dims, k):
    o = gen(*dims)
    d = to_dev(ctx, o)
    x = np.random.default_rng(3).standard_normal((o.nrows, k))
    l1 = O.new_l1(o).reshape(-1, 1)
    got = F.ErrorPropogator(F.SparseMatOp(d), F.new_l1(d)).apply(x)
    want = O.error_propagator(o, lambda r: l1 * r, x)  # adaptivity.rs:191-198
    if d.plan()["threads_per_row"] == 1:
        assert np.array_equal(got, want)
    else:
        assert np.max(np.abs(got - want)) < 1e-12 * np.max(np.abs(x)) * 30
    # non-diagonal preconditioner: SpMM, block-smoother apply, axpby
    part, _ = F.partitioners.geometric_partition(dims)
    bs = F.BlockSmoother.new(F.SparseMatOp(d), part)
    got = F.ErrorPropogator(F.SparseMatOp(d), bs).apply(x)
    want = O.error_propagator(o, lambda r: O.block_smoother_apply(o, part.agg_ptr, part.agg_nodes, r), x)
    assert np.max(np.abs(got - want)) < 1e-11 * np.max(np.abs(x))


@pytest.mark.parametrize("gen,dims,k,iters", [(O.gen_g7, (10, 9, 8), 4, 6), (O.gen_g27, (7, 6, 5), 6, 4), (O.gen_g7, (12, 12, 12), 1, 10)])
def test_smooth_vector_matches_oracle(ctx, F, gen, dims, k, iters):
    o = gen(*dims)
    d = to_dev(ctx, o)
    x0 = np.random.default_rng(11).standard_normal((o.nrows, k))
    l1 = O.new_l1(o).reshape(-1, 1)
    want, want_cfs = O.smooth_vector(o, lambda r: l1 * r, x0, iters)  # adaptivity.rs:307-390
    got, cfs = F.adaptivity.smooth_vector(F.SparseMatOp(d), F.new_l1(d), iters, k, x0=x0)
    assert np.max(np.abs(got.T @ got - np.eye(k))) < 1e-13
    assert np.max(np.abs(got - want)) < 1e-9
    assert np.allclose(cfs, want_cfs, rtol=1e-9, atol=0)
    assert all(0.0 < c < 1.0 for c in cfs)  # E is an A-norm contraction for the L1 smoother
    # iterations = 0: two QRs only; cfs still reported
    got0, _ = F.adaptivity.smooth_vector(F.SparseMatOp(d), F.new_l1(d), 0, k, x0=x0)
    assert np.max(np.abs(got0 - O.thin_q(O.thin_q(x0)))) < 1e-12


def test_smooth_vector_block_preconditioner(ctx, F):
    dims = (8, 8, 6)
    o = O.gen_g7(*dims)
    d = to_dev(ctx, o)
    part, _ = F.partitioners.geometric_partition(dims)
    bs = F.BlockSmoother.new(F.SparseMatOp(d), part)
    x0 = np.random.default_rng(12).standard_normal((o.nrows, 3))
    want, want_cfs = O.smooth_vector(o, lambda r: O.block_smoother_apply(o, part.agg_ptr, part.agg_nodes, r), x0, 5)
    got, cfs = F.adaptivity.smooth_vector(F.SparseMatOp(d), bs, 5, 3, x0=x0)
    assert np.max(np.abs(got - want)) < 1e-9 and np.allclose(cfs, want_cfs, rtol=1e-8, atol=0)


def test_find_near_null_with_algebraic_partitioner(ctx, F):
    """adaptivity.rs:264-305 end to end: L1 search, create_weights, algebraic block smoother, search."""
    dims = (10, 10, 8)
    d = F.gallery.poisson7(ctx, *dims)
    op = F.SparseMatOp(d)
    basis = F.adaptivity.find_near_null(op, 8, 4, 16.0, seed=7)
    assert basis.shape == (d.nrows, 4) and np.max(np.abs(basis.T @ basis - np.eye(4))) < 1e-12
    # the smoothest Laplacian mode dominates the span: Rayleigh quotients far below the spectrum's middle
    a = d.to_scipy()
    rq = np.array([basis[:, c] @ (a @ basis[:, c]) for c in range(4)])
    assert np.all(rq < 0.5 * 6.0)


def test_hierarchy_from_algebraic_partitioner_matches_oracle(ctx, F):
    """AggregationConfig with a PartitionerConfig (interpolation/mod.rs:129-156): aggregates, P, R, A_c
    equal the oracle's built from the restated partitioner; PCG iteration counts equal (+-1).
    Two levels only: with a single near-null vector the LS distances are 0 up to rounding
    (rho^2 = 1), so from level 1 on the strength graph is decided by the last bits of the coarse
    near-null and is not comparable between two implementations (nor between two runs of the
    reference)."""
    dims = (10, 8, 6)
    o = O.gen_g27(*dims, ey=1.0, ez=1e-2)
    d = to_dev(ctx, o)
    n = o.nrows
    nn = np.full((n, 1), 1.0 / np.sqrt(n))
    weights = O.create_weights(o, nn)
    pcfg = F.partitioners.PartitionerConfig(8.0, 1.0, 30)
    cfg = F.HierarchyConfig(coarsest_dim=100, interpolation_config=F.AggregationConfig(1, 1, pcfg))
    h = cfg.build(F.SparseMatOp(d), nn, weights)

    def orc_partitioner(level, fine, near_null):
        n2a, _ = OP.build_partition(fine.row_ptr, fine.col, near_null, weights, 8.0, 1.0, 30)
        return n2a

    ho = O.build_hierarchy(o, nn, dims, coarsest_dim=100, partitioner=orc_partitioner)
    assert h.levels() == ho.levels == 2
    assert np.array_equal(h.get_partition(0).agg_ptr, ho.partitions[0][0])
    assert np.array_equal(h.get_partition(0).agg_nodes, ho.partitions[0][1])
    assert O.mats_are_equal(to_oracle(h.get_interpolation(0)), ho.interpolations[0])
    assert O.mats_are_equal(to_oracle(h.get_restriction(0)), ho.restrictions[0])
    assert O.mats_are_equal(to_oracle(h.get_arc_mat(1)), ho.operators[1])
    b = np.ones(n)
    mg = F.MultigridConfig(smoother="l1").build(h)
    x = np.zeros(n)
    info = F.conjugate_gradient(x, mg, d, b, F.CgParams(0.0, 1e-8, 200))
    _, oinfo = O.pcg(o, b, O.multigrid_from_hierarchy(ho, "l1"), rel_tol=1e-8, max_iters=200)
    assert abs(info.iter_count - oinfo.iters) <= 1
    assert np.linalg.norm(b - O.spmm_csr(o, x).ravel()) < 1e-8 * np.linalg.norm(b) * 1.01


def test_algebraic_hierarchy_multilevel_converges(ctx, F):
    """Anisotropic 27-point operator, two near-null candidates, algebraic aggregates on every level:
    valid partitions, shrinking levels, PCG + V(1,1) converges; the strength-aware aggregates need
    fewer iterations than geometric 2x2x2 boxes on the same operator."""
    dims = (16, 16, 16)
    a = F.gallery.diffusion27(ctx, *dims, 1.0, 1e-2)
    n = a.nrows
    op = F.SparseMatOp(a)
    nn = np.full((n, 1), 1.0 / np.sqrt(n))
    weights = F.adaptivity.create_weights(nn, a)
    cfg = F.HierarchyConfig(coarsest_dim=60, interpolation_config=F.AggregationConfig(1, 1, F.PartitionerConfig(8.0, 1.0, 30)))
    h = cfg.build(op, nn, weights)
    assert h.levels() >= 3
    rows = [m.mat_ref().nrows for m in h.operators()]
    assert all(r1 < r0 for r0, r1 in zip(rows, rows[1:]))
    for lvl in range(h.levels() - 1):
        h.get_partition(lvl).validate()
        assert h.get_partition(lvl).nnodes() == rows[lvl] and h.get_partition(lvl).naggs() == rows[lvl + 1]
    b, x = np.ones(n), np.zeros(n)
    info = F.conjugate_gradient(x, F.MultigridConfig(smoother="l1").build(h), a, b, F.CgParams(0.0, 1e-8, 300))
    hg = F.HierarchyConfig(60, F.AggregationConfig(1, 1, F.GeometricPartitioner(dims))).build(op, nn)
    xg = np.zeros(n)
    ig = F.conjugate_gradient(xg, F.MultigridConfig(smoother="l1").build(hg), a, b, F.CgParams(0.0, 1e-8, 300))
    assert info.rel_residual < 1e-8 and ig.rel_residual < 1e-8
    assert np.linalg.norm(x - xg) < 1e-6 * np.linalg.norm(xg)
    print(f"algebraic: {info.iter_count} iterations, levels {rows}; geometric: {ig.iter_count}")
