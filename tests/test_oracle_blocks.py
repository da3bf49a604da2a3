"""CPU checks that pin the *oracle's* restatements added for SURVEY 8a18 / 8f against independent dense
formulas (the reference ships no vectors for them): block_jacobi, smooth_p, the vector block smoother,
Composite, smooth_vector, create_weights."""
import numpy as np
import scipy.sparse as sp

import oracle as O


def vector_operator(dims, vdim, seed=0):
    rng = np.random.default_rng(seed)
    g = O.gen_g7(*dims).to_scipy()
    q = rng.standard_normal((vdim, vdim))
    b = q @ q.T + vdim * np.eye(vdim)
    a = sp.kron(g, b).tolil()
    for node in range(g.shape[0]):
        d = rng.standard_normal((vdim, vdim))
        a[node * vdim:(node + 1) * vdim, node * vdim:(node + 1) * vdim] += 0.1 * (d + d.T) + np.eye(vdim)
    a = a.tocsr()
    a.sort_indices()
    return O.Csr.from_scipy(a)


def test_block_jacobi_and_smooth_p_equal_their_dense_formulas():
    vdim, dims = 3, (4, 3, 2)
    o = vector_operator(dims, vdim)
    a = o.to_scipy().toarray()
    n = o.nrows
    nn = np.random.default_rng(1).standard_normal((n, 2))
    ap, an, _ = O.geometric_aggregates(dims)
    p, _ = O.tentative_p(n, nn, ap, an, 2, vdim)
    pd = p.to_scipy().toarray()
    dinv = np.zeros((n, n))
    for b in range(n // vdim):
        s = slice(b * vdim, (b + 1) * vdim)
        dinv[s, s] = np.linalg.inv(a[s, s])
    want = pd - 0.66 * dinv @ a @ pd                                   # interpolation/mod.rs:963-1028
    got = O.block_jacobi(o, vdim, p).to_scipy().toarray()
    assert np.max(np.abs(got - want)) < 1e-13 * np.max(np.abs(want))
    minv = O.block_diag_inverse(o, vdim, 1.0)
    assert np.max(np.abs(minv.to_scipy().toarray() - dinv)) < 1e-14
    got = O.smooth_p(o, minv, p).to_scipy().toarray()                  # :1030-1040
    assert np.max(np.abs(got - (pd - dinv @ a @ pd))) < 1e-13 * np.max(np.abs(pd))
    # the pattern is the structural product's (explicit zeros kept), containing P's
    s = O.block_jacobi(o, vdim, p)
    assert s.nnz == O.spgemm(O.block_diag_inverse(o, vdim, 1.0), O.spgemm(o, p)).nnz >= p.nnz


def test_vector_block_smoother_is_the_compensated_block_inverse():
    vdim, dims = 2, (4, 4, 2)
    o = vector_operator(dims, vdim, seed=2)
    a = o.to_scipy().toarray()
    ap, an, _ = O.geometric_aggregates(dims)
    n = o.nrows
    m = np.zeros((n, n))
    for g in range(len(ap) - 1):
        nodes = an[ap[g]:ap[g + 1]]
        dofs = (nodes[:, None] * vdim + np.arange(vdim)).reshape(-1)
        blk = a[np.ix_(dofs, dofs)].copy()
        inside = set(int(x) for x in nodes)
        for li, bi in enumerate(nodes):
            rows = slice(bi * vdim, (bi + 1) * vdim)
            for bj in range(n // vdim):
                if bj in inside:
                    continue
                aij = a[rows, bj * vdim:(bj + 1) * vdim]
                if not np.any(aij):
                    continue
                # 0.5 * sqrt(A_ij A_ij^T): the symmetric polar factor, independent of SVD sign choices
                w, v = np.linalg.eigh(aij @ aij.T)
                blk[li * vdim:(li + 1) * vdim, li * vdim:(li + 1) * vdim] += 0.5 * (v * np.sqrt(np.maximum(w, 0))) @ v.T
        m[np.ix_(dofs, dofs)] = np.linalg.inv(blk)
    r = np.random.default_rng(3).standard_normal((n, 2))
    got = O.block_smoother_vector_apply(o, vdim, ap, an, r)
    assert np.max(np.abs(got - m @ r)) < 1e-9 * np.max(np.abs(m @ r))
    assert np.all(np.linalg.eigvalsh(0.5 * (m + m.T)) > 0)


def test_composite_is_the_symmetric_multiplicative_combination():
    o = O.gen_g7(4, 4, 3)
    a = o.to_scipy().toarray()
    n = o.nrows
    d1 = O.new_l1(o).reshape(-1, 1)
    d2 = O.new_jacobi(o, 0.5).reshape(-1, 1)
    rng = np.random.default_rng(4)
    rhs = rng.standard_normal((n, 2))
    eye = np.eye(n)
    e1, e2 = eye - np.diagflat(d1) @ a, eye - np.diagflat(d2) @ a
    # composite.rs:66-83 with components [c1, c2]: reversed pass applies c2 then c1, the forward pass skips
    # the first component and applies c2 again: error propagator E2 E1 E2  ->  B = (I - E2 E1 E2) A^-1
    want = (eye - e2 @ e1 @ e2) @ np.linalg.solve(a, rhs)
    got = O.composite_apply(o, [lambda r: d1 * r, lambda r: d2 * r], rhs)
    assert np.max(np.abs(got - want)) < 1e-12 * np.max(np.abs(want))
    one = O.composite_apply(o, [lambda r: d1 * r], rhs)
    assert np.array_equal(one, 0.0 + d1 * rhs)
    assert np.array_equal(O.composite_apply(o, [], rhs), np.zeros_like(rhs))


def test_smooth_vector_and_weights():
    o = O.gen_g7(5, 4, 3)
    a = o.to_scipy().toarray()
    n = o.nrows
    d = O.new_l1(o).reshape(-1, 1)
    x0 = np.random.default_rng(5).standard_normal((n, 3))
    x, cfs = O.smooth_vector(o, lambda r: d * r, x0, 7)
    assert np.max(np.abs(x.T @ x - np.eye(3))) < 1e-13
    # span(x) = span(E^7 x0) for E = I - D A (QR steps do not change the subspace)
    e = np.eye(n) - np.diagflat(d) @ a
    y = np.linalg.matrix_power(e, 7) @ x0
    q, _ = np.linalg.qr(y)
    assert np.allclose(np.linalg.svd(q.T @ x, compute_uv=False), 1.0, atol=1e-10)
    for c in range(3):
        w = x[:, c]
        ew = e @ w
        assert np.isclose(cfs[c], np.sqrt(ew @ a @ ew) / np.sqrt(w @ a @ w), rtol=1e-12)
    wts = O.create_weights(o, x)
    assert np.allclose(wts, [1.0 / (x[:, c] @ a @ x[:, c]) for c in range(3)], rtol=1e-13)
