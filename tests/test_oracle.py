"""CPU tests: the oracle against scipy, analytic identities and the committed golden fixtures.
(The reference has no tests of its own; 8c lists these substitutes.)"""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

import oracle as O

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.json")))
EPS = float(np.finfo(float).eps)


def g7_scipy(nx, ny, nz):
    def K(n):
        return sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1])
    I = sp.identity
    return (sp.kron(I(nz), sp.kron(I(ny), K(nx))) + sp.kron(I(nz), sp.kron(K(ny), I(nx)))
            + sp.kron(K(nz), sp.kron(I(ny), I(nx)))).tocsr()


def g27_scipy(n, ey, ez):
    K = sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1])
    M = sp.diags([np.ones(n - 1) / 6, 4 * np.ones(n) / 6, np.ones(n - 1) / 6], [-1, 0, 1])
    return (sp.kron(M, sp.kron(M, K)) + ey * sp.kron(M, sp.kron(K, M)) + ez * sp.kron(K, sp.kron(M, M))).tocsr()


def test_generators_match_scipy():
    a = O.gen_g7(6, 5, 4)
    s = g7_scipy(6, 5, 4); s.sort_indices()
    assert a.nnz == 7 * 120 - 2 * (30 + 20 + 24)
    assert np.array_equal(a.col, s.indices) and np.array_equal(a.val, s.data)
    b = O.gen_g27(8)
    t = g27_scipy(8, 1.0, 1e-2); t.sort_indices()
    assert b.nnz == (3 * 8 - 2) ** 3
    assert np.array_equal(b.col, t.indices) and np.allclose(b.val, t.data, rtol=1e-14, atol=0)
    w = np.linalg.eigvalsh(t.toarray()[:64, :64])
    assert w.min() > 0  # SPD principal block


def test_g7_row_sums_vanish_inside():
    n = 8
    a = O.gen_g7(n)
    y = O.spmm_csr(a, np.ones(a.nrows)).reshape(n, n, n)
    assert np.all(y[1:-1, 1:-1, 1:-1] == 0) and np.all(y[0] > 0)


def test_triplets_sum_duplicates_keep_zeros():
    m = O.Csr.from_triplets(3, 4, [2, 0, 0, 2, 1], [1, 3, 3, 0, 2], [1.0, 2.0, 3.0, 0.0, 5.0])
    assert m.row_ptr.tolist() == [0, 1, 2, 4]
    assert m.col.tolist() == [3, 2, 0, 1] and m.val.tolist() == [5.0, 5.0, 0.0, 1.0]


def test_parspmm_equals_serial_csr():
    rng = np.random.default_rng(0)
    a = O.gen_g7(30, 29, 28)  # 24360 rows: three 8192-row block rows, last one ragged
    x = rng.standard_normal((a.nrows, 3))
    op = O.ParSpmmOp(a, 4)
    assert np.array_equal(op.apply(x), O.spmm_csr(a, x))  # same ascending-column summation order
    with pytest.raises(ValueError):
        O.ParSpmmOp(a, 1)  # par_spmm.rs:33-36


def test_spgemm_transpose_vs_scipy():
    rng = np.random.default_rng(1)
    a = sp.random(60, 50, density=0.1, random_state=2, format="csr")
    b = sp.random(50, 40, density=0.15, random_state=3, format="csr")
    c = O.spgemm(O.Csr.from_scipy(a), O.Csr.from_scipy(b))
    ref = (a @ b).tocsr(); ref.sort_indices()
    structural = ((abs(a) > 0).astype(float) @ (abs(b) > 0).astype(float)).tocsr(); structural.sort_indices()
    assert np.array_equal(c.row_ptr, structural.indptr) and np.array_equal(c.col, structural.indices)
    assert abs(c.to_scipy() - ref).max() < 1e-14
    t = O.transpose(O.Csr.from_scipy(a))
    tt = a.T.tocsr(); tt.sort_indices()
    assert np.array_equal(t.col, tt.indices) and np.array_equal(t.val, tt.data)


def test_unsmoothed_aggregation_is_half_coarse_laplacian():
    ap, an, dc = O.geometric_aggregates((8, 8, 8))
    a = O.gen_g7(8)
    g = O.smoothed_aggregation(a, ap, an, np.full((512, 1), 1 / np.sqrt(512)), 1, 0)
    half = O.gen_g7(4)
    assert np.array_equal(g.coarse_mat.col, half.col)
    assert np.max(np.abs(g.coarse_mat.val - 0.5 * half.val)) < 1e-15
    ptp = O.spgemm(g.restriction, g.interpolation)
    assert ptp.nnz == 64 and np.allclose(ptp.val, 1.0)  # P^T P = I for the tentative P


def test_smoothed_rap_matches_dense():
    n = 6
    a = O.gen_g7(n)
    ap, an, _ = O.geometric_aggregates((n, n, n))
    rng = np.random.default_rng(5)
    nn = 1.0 + 0.1 * rng.standard_normal((a.nrows, 1))
    g = O.smoothed_aggregation(a, ap, an, nn, 1, 1)
    A = a.to_scipy().toarray(); P = g.interpolation.to_scipy().toarray()
    assert np.allclose(g.coarse_mat.to_scipy().toarray(), P.T @ A @ P, rtol=0, atol=1e-13)
    ac = g.coarse_mat.to_scipy()
    assert abs(ac - ac.T).max() < 1e-14
    assert np.array_equal(g.restriction.to_scipy().toarray(), P.T)


def test_tentative_p_multi_candidate_orthonormal():
    rng = np.random.default_rng(7)
    n = 4
    ap, an, _ = O.geometric_aggregates((n, n, n))
    nn = rng.standard_normal((n ** 3, 3))
    p, cnn = O.tentative_p(n ** 3, nn, ap, an, cand=2)
    P = p.to_scipy().toarray()
    assert np.allclose(P.T @ P, np.eye(P.shape[1]), atol=1e-13)
    # P * coarse_nn reproduces the rank-2 truncation of every aggregate block
    for a in range(len(ap) - 1):
        nodes = an[ap[a]:ap[a + 1]]
        u, s, vt = np.linalg.svd(nn[nodes], full_matrices=False)
        assert np.allclose(P[nodes][:, 2 * a:2 * a + 2] @ cnn[2 * a:2 * a + 2], (u[:, :2] * s[:2]) @ vt[:2], atol=1e-12)


def test_multigrid_is_symmetric():
    """symmetry_test (multigrid.rs:520-580): u^T (B v) == v^T (B u)."""
    a = O.gen_g7(8)
    h = O.build_hierarchy(a, np.full((512, 1), 1 / np.sqrt(512)), (8, 8, 8), coarsest_dim=100)
    mg = O.multigrid_from_hierarchy(h, "l1")
    rng = np.random.default_rng(11)
    u, v = rng.standard_normal((512, 4)), rng.standard_normal((512, 4))
    utbv, vtbu = u.T @ mg.apply(v), v.T @ mg.apply(u)
    assert np.max(np.abs(utbv - vtbu.T)) < 1e-12 * np.max(np.abs(utbv))


def test_simple_geometric_table_matches_golden():
    """BASELINE config #1 (examples/simple_geometric.rs): mesh-independent PCG+MG counts."""
    rows = GOLD["simple_geometric"]
    for row in rows[:5]:
        ne = row["dofs"] + 1
        refinement = int(np.log2(ne // 10))
        a = O.gen_g1(ne)
        mg = O.Multigrid()
        mg.add_level(a, O.new_jacobi(a, 0.66))
        for level in range(1, refinement + 1):
            ce = 10 * 2 ** (refinement - level)
            m = O.gen_g1(ce)
            mg.add_level(m, "cholesky" if level == refinement else O.new_jacobi(m, 0.66), O.gen_g1_restrict(ce - 1),
                         O.gen_g1_interp(ce - 1))
        b = np.ones(ne - 1)
        _, i1 = O.pcg(a, b, O.new_jacobi(a, 0.66), rel_tol=1e-8, abs_tol=EPS, max_iters=6000)
        _, i2 = O.pcg(a, b, mg, rel_tol=1e-8, abs_tol=EPS, max_iters=6000)
        _, i3 = O.stationary_solver(a, b, mg, 6000, 1e-8)
        assert (i1.iters, i2.iters, i3) == (row["pcg_jacobi"], row["pcg_mg"], row["stat_mg"])
    mg_counts = [r["pcg_mg"] for r in rows]
    assert max(mg_counts) - min(mg_counts) <= 2  # mesh independence (simple_geometric.rs:50-51)
    assert rows[-1]["pcg_jacobi"] > 50 * rows[-1]["pcg_mg"]


def test_pcg_against_scipy_solution():
    a = O.gen_g7(10)
    b = np.ones(a.nrows)
    x, info = O.pcg(a, b, O.new_jacobi(a, 1.0), rel_tol=1e-12)
    from scipy.sparse.linalg import spsolve
    assert info.status == 0 and np.allclose(x, spsolve(a.to_scipy().tocsc(), b), rtol=1e-9)
    _, capped = O.pcg(a, b, None, rel_tol=1e-14, max_iters=3)
    assert capped.status == 1 and capped.iters == 3  # CgError::NoConvergence


def test_amg_golden_small():
    g = GOLD["amg"]["g7_16_l1"]
    a = O.gen_g7(16)
    h = O.build_hierarchy(a, np.full((a.nrows, 1), 1 / np.sqrt(a.nrows)), (16, 16, 16))
    assert [o.nrows for o in h.operators] == g["level_rows"] and [o.nnz for o in h.operators] == g["level_nnz"]
    mg = O.multigrid_from_hierarchy(h, "l1")
    _, info = O.pcg(a, np.ones(a.nrows), mg, rel_tol=1e-8)
    assert info.iters == g["iters"]["1e-08"]["iters"]


def test_block_smoother_matches_dense_blocks():
    a = O.gen_g7(4)
    ap, an, _ = O.geometric_aggregates((4, 4, 4))
    r = np.random.default_rng(3).standard_normal((64, 2))
    out = O.block_smoother_apply(a, ap, an, r)
    A = a.to_scipy().toarray()
    d = np.diag(A)
    for g in range(len(ap) - 1):
        nodes = an[ap[g]:ap[g + 1]]
        blk = A[np.ix_(nodes, nodes)].copy()
        for li, i in enumerate(nodes):
            for j in np.nonzero(A[i])[0]:
                if j not in nodes:
                    blk[li, li] += 0.5 * np.sqrt(d[i] / d[j]) * abs(A[i, j])
        assert np.allclose(out[nodes], np.linalg.solve(blk, r[nodes]), rtol=1e-12)


# ---------------------------------------------------------------------- second, independent restatement
def _numpy_cycle(levels, mu, nu, v, f, lvl):
    """multigrid.rs:269-380 written directly in dense numpy (no shared code with the C oracle)."""
    A, minv = levels[lvl][0], levels[lvl][1]
    if lvl == len(levels) - 1:
        return minv(f)
    R, P = levels[lvl + 1][2], levels[lvl + 1][3]  # transfer operators are stored with the coarse level
    for _ in range(nu):
        v = v + minv(f - A @ v)
    fc = R @ (f - A @ v)
    vc = np.zeros((levels[lvl + 1][0].shape[0], f.shape[1]))
    for _ in range(mu):
        vc = _numpy_cycle(levels, mu, nu, vc, fc, lvl + 1)
    v = v + P @ vc
    for _ in range(nu):
        v = v + minv(f - A @ v)
    return v


@pytest.mark.parametrize("mu,nu", [(1, 1), (2, 1), (1, 3)])
def test_cycle_against_dense_numpy_restatement(mu, nu):
    dims = (8, 8, 4)
    a = O.gen_g7(*dims)
    n = a.nrows
    h = O.build_hierarchy(a, np.full((n, 1), 1 / np.sqrt(n)), dims, coarsest_dim=20)
    assert h.levels == 3
    mg = O.multigrid_from_hierarchy(h, "l1", mu=mu, nu=nu)
    levels = []
    for lvl, op in enumerate(h.operators):
        A = op.to_scipy().toarray()
        if lvl == h.levels - 1:
            minv = (lambda Ad: (lambda r: np.linalg.solve(Ad, r)))(A)
        else:
            d = 1.0 / np.abs(A).sum(axis=1)
            minv = (lambda dd: (lambda r: dd[:, None] * r))(d)
        R = h.restrictions[lvl - 1].to_scipy().toarray() if lvl else None
        P = h.interpolations[lvl - 1].to_scipy().toarray() if lvl else None
        levels.append((A, minv, R, P))
    f = np.random.default_rng(17).standard_normal((n, 2))
    want = _numpy_cycle(levels, mu, nu, np.zeros_like(f), f, 0)
    got = mg.apply(f)
    assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))


def test_pcg_against_numpy_restatement():
    """Textbook PCG with faer's stopping rule, re-written in numpy: same iteration count and iterates."""
    a = O.gen_g27(6, 5, 4)
    A = a.to_scipy().tocsr()
    d = O.new_jacobi(a, 1.0)
    b = np.random.default_rng(3).standard_normal(a.nrows)
    x = np.zeros_like(b); r = b.copy(); thr = 1e-10 * np.linalg.norm(b)
    z = d * r; p = z.copy(); rtz = r @ z; it = 0
    while True:
        q = A @ p
        alpha = rtz / (p @ q)
        x += alpha * p; r -= alpha * q; it += 1
        if np.linalg.norm(r) < thr or it >= 1000:
            break
        z = d * r; rtz_new = r @ z; p = z + (rtz_new / rtz) * p; rtz = rtz_new
    xo, info = O.pcg(a, b, d, rel_tol=1e-10, abs_tol=0.0, max_iters=1000)
    assert info.status == 0 and info.iters == it
    assert np.linalg.norm(xo - x) <= 1e-12 * np.linalg.norm(x)
