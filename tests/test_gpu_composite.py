"""GPU parity tests of the Composite preconditioner (src/preconditioners/composite.rs:66-83) and of
the adaptive driver built on it (adaptivity.rs:28-165), through the C ABI against the oracle."""
import numpy as np
import pytest

import oracle as O
from util import to_dev

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    import faer_amg_b200 as F
    return F


def _two_multigrids(ctx, F, dims):
    o = O.gen_g7(*dims)
    d = to_dev(ctx, o)
    n = o.nrows
    nn1 = np.full((n, 1), 1.0 / np.sqrt(n))
    x = np.arange(n) % dims[0]
    nn2 = O.thin_q((1.0 + 0.3 * np.sin(2 * np.pi * (x + 0.5) / dims[0])).reshape(-1, 1))
    mgs, omgs = [], []
    for nn, sm in ((nn1, "l1"), (nn2, "jacobi")):
        h = F.HierarchyConfig(40, F.AggregationConfig(1, 1, F.GeometricPartitioner(dims))).build(F.SparseMatOp(d), nn)
        mgs.append(F.MultigridConfig(smoother=sm).build(h))
        omgs.append(O.multigrid_from_hierarchy(O.build_hierarchy(o, nn, dims, coarsest_dim=40), sm))
    return o, d, mgs, omgs


def test_composite_apply_matches_oracle(ctx, F):
    dims = (12, 10, 8)
    o, d, mgs, omgs = _two_multigrids(ctx, F, dims)
    l1 = F.new_l1(d)
    ol1 = O.new_l1(o).reshape(-1, 1)
    rhs = np.random.default_rng(1).standard_normal((o.nrows, 3))
    comps = [mgs[0], l1, mgs[1]]
    ocomps = [omgs[0].apply, lambda r: ol1 * r, omgs[1].apply]
    for m in (1, 2, 3):
        c = F.Composite.new_with_components(d, comps[:m])
        got = c.apply(rhs)
        want = O.composite_apply(o, ocomps[:m], rhs)
        assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want)), m
    # one component: the composite is that component (composite.rs:71-76 with an empty second loop)
    assert np.array_equal(F.Composite.new(d, mgs[0]).apply(rhs[:, :1]), mgs[0].apply(rhs[:, :1]))
    # symmetric: <u, C v> == <C u, v>
    c = F.Composite.new_with_components(d, comps)
    u, v = rhs[:, :1], rhs[:, 1:2]
    assert (u.T @ c.apply(v)).item() == pytest.approx((c.apply(u).T @ v).item(), rel=1e-10)
    bad = F.Composite.new(d, F.new_l1(to_dev(ctx, O.gen_g7(3))))  # a component of the wrong dimension
    with pytest.raises(F.FamgError):
        bad.apply(rhs)


def test_pcg_with_composite_matches_oracle_counts(ctx, F):
    dims = (12, 10, 8)
    o, d, mgs, omgs = _two_multigrids(ctx, F, dims)
    b = np.ones(o.nrows)
    x1, x2 = np.zeros(o.nrows), np.zeros(o.nrows)
    i1 = F.conjugate_gradient(x1, mgs[0], d, b, F.CgParams(0.0, 1e-10, 200)).iter_count
    comp = F.Composite.new_with_components(d, mgs)
    i2 = F.conjugate_gradient(x2, comp, d, b, F.CgParams(0.0, 1e-10, 200)).iter_count
    assert i2 < i1 and np.linalg.norm(x1 - x2) < 1e-8 * np.linalg.norm(x1)
    # numpy PCG with the oracle composite as the preconditioner (faer's stopping rule)
    a = o.to_scipy()
    x = np.zeros(o.nrows); r = b.copy(); thr = 1e-10 * np.linalg.norm(b)
    z = O.composite_apply(o, [m.apply for m in omgs], r)[:, 0]; p = z.copy(); rz = r @ z; it = 0
    while np.linalg.norm(r) >= thr and it < 200:
        q = a @ p; alpha = rz / (p @ q); x += alpha * p; r -= alpha * q; it += 1
        if np.linalg.norm(r) < thr:
            break
        z = O.composite_apply(o, [m.apply for m in omgs], r)[:, 0]; rz_new = r @ z; p = z + (rz_new / rz) * p; rz = rz_new
    assert abs(i2 - it) <= 1


def test_smooth_vector_with_multigrid_and_composite(ctx, F):
    dims = (10, 8, 6)
    o, d, mgs, omgs = _two_multigrids(ctx, F, dims)
    x0 = np.random.default_rng(3).standard_normal((o.nrows, 3))
    want, wcfs = O.smooth_vector(o, omgs[0].apply, x0, 4)
    got, cfs = F.smooth_vector(F.SparseMatOp(d), mgs[0], 4, 3, x0=x0)
    # the multigrid error propagator damps everything but a thin subspace: the iterates are compared as
    # subspaces (principal angles) and through their convergence factors
    assert np.allclose(np.linalg.svd(want.T @ got, compute_uv=False), 1.0, atol=1e-6)
    assert np.allclose(cfs, wcfs, rtol=1e-6)
    comp = F.Composite.new_with_components(d, mgs)
    got2, cfs2 = F.smooth_vector(F.SparseMatOp(d), comp, 3, 3, x0=x0)
    want2, wcfs2 = O.smooth_vector(o, lambda r: O.composite_apply(o, [m.apply for m in omgs], r), x0, 3)
    assert np.allclose(cfs2, wcfs2, rtol=1e-5) and max(cfs2) < max(cfs)


def test_adaptive_config_builds_a_composite(ctx, F):
    """AdaptiveConfig::build (adaptivity.rs:54-165) end to end on a small anisotropic operator."""
    dims = (12, 12, 8)
    a = F.gallery.diffusion27(ctx, *dims, 1.0, 1e-2)
    op = F.SparseMatOp(a)
    hcfg = F.HierarchyConfig(60, F.AggregationConfig(1, 2, F.PartitionerConfig(8.0, 1.0, 20)))
    cfg = F.AdaptiveConfig(hcfg, F.MultigridConfig(smoother="l1"), max_components=2, test_iters=6, coarsening_near_null_dim=4,
                           smoothing_block_size=8.0, seed=11)
    comp = cfg.build(op)
    assert len(comp.components()) == 2
    b, x = np.ones(a.nrows), np.zeros(a.nrows)
    info = F.conjugate_gradient(x, comp, a, b, F.CgParams(0.0, 1e-8, 300))
    assert info.rel_residual < 1e-8
    x1 = np.zeros(a.nrows)
    i1 = F.conjugate_gradient(x1, comp.components()[0], a, b, F.CgParams(0.0, 1e-8, 300))
    assert info.iter_count <= i1.iter_count and np.linalg.norm(x - x1) < 1e-6 * np.linalg.norm(x1)
