"""GPU parity tests: every kernel of the hot path, called through the C ABI, against the CPU oracle
on the same inputs.  Bar (north-star): bit-exact sparsity patterns and integer/index results;
f64 values within 1e-12 relative -- and in fact bit-identical wherever the GPU keeps the
reference's summation order (threads-per-row == 1, SpGEMM, diagonal setup)."""
import json
import os

import numpy as np
import pytest

import oracle as O
from util import assert_rel, random_csr, same_pattern, spmv_bound, to_dev, to_oracle

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.json")))
EPS = float(np.finfo(float).eps)


@pytest.fixture(scope="module")
def F():
    import faer_amg_b200 as F
    return F


@pytest.fixture(params=[1, 2], ids=["staged-cta", "tma-pipeline"])
def variant(request, ctx):
    """Run a test under both SpMV kernel variants (the persistent TMA pipeline is normally reserved
    for operators with >= 131072 rows)."""
    ctx.set_option("spmv_variant", request.param)
    ctx.set_option("tma_min_rows", 1)
    yield request.param
    ctx.set_option("spmv_variant", 2)
    ctx.set_option("tma_min_rows", 1 << 17)


# ------------------------------------------------------------------ containers / generators
def test_gallery_bit_exact(ctx, F):
    for dims in [(8, 8, 8), (7, 5, 3), (33, 2, 1)]:
        a, o = F.gallery.poisson7(ctx, *dims), O.gen_g7(*dims)
        assert same_pattern(a, o) and np.array_equal(a.to_host()[2], o.val)
        b, p = F.gallery.diffusion27(ctx, *dims), O.gen_g27(*dims)
        assert same_pattern(b, p) and np.array_equal(b.to_host()[2], p.val)


def test_csr_roundtrip_and_triplets(ctx, F):
    rng = np.random.default_rng(0)
    o = random_csr(rng, 200, 150, rng.integers(0, 12, 200))
    d = to_dev(ctx, o)
    assert same_pattern(d, o) and np.array_equal(d.to_host()[2], o.val)
    r, c, v = rng.integers(0, 40, 500), rng.integers(0, 30, 500), rng.standard_normal(500)
    v[::7] = 0.0
    t = F.SparseRowMat.try_new_from_triplets(ctx, 40, 30, r, c, v)
    ot = O.Csr.from_triplets(40, 30, r, c, v)
    assert same_pattern(t, ot) and np.array_equal(t.to_host()[2], ot.val)
    with pytest.raises(F.FamgError):
        F.SparseRowMat.from_csr(ctx, 2, 2, [0, 2, 2], [1, 0], [1.0, 2.0])  # unsorted row
    with pytest.raises(ValueError):
        F.SparseMatOp(to_dev(ctx, random_csr(rng, 4, 5, [1] * 4)))  # core.rs:57-59 square only


# ------------------------------------------------------------------ SpMV / SpMM family
CASES = [("g7", (24, 20, 18)), ("g27", (12, 11, 10)), ("g1", 999)]


def _case(name, arg):
    if name == "g7":
        return O.gen_g7(*arg)
    if name == "g27":
        return O.gen_g27(*arg)
    return O.gen_g1(arg + 1)


@pytest.mark.parametrize("name,arg", CASES)
@pytest.mark.parametrize("k", [1, 3])
def test_spmm_matches_oracle(ctx, F, variant, name, arg, k):
    o = _case(name, arg)
    d = to_dev(ctx, o)
    x = np.random.default_rng(1).standard_normal((o.ncols, k))
    got, want = d.apply(x), O.spmm_csr(o, x)
    if d.plan()["threads_per_row"] == 1:
        assert np.array_equal(got, want)  # same ascending-column summation order as the reference
    assert_rel(got, want, spmv_bound(o, x))
    par = O.ParSpmmOp(o, 4).apply(x)  # the reference's tiled operator gives the same numbers
    assert_rel(got, par, spmv_bound(o, x))


def test_spmm_irregular_rows_and_fallback_path(ctx, F, variant):
    """Empty rows, ragged rows, rows longer than one CTA's staging buffer (direct path), rectangular."""
    rng = np.random.default_rng(2)
    lens = rng.integers(0, 9, 700)
    lens[5] = 0; lens[6] = 0; lens[100] = 3000; lens[101] = 2800; lens[699] = 1500
    o = random_csr(rng, 700, 4000, lens)
    d = to_dev(ctx, o)
    x = rng.standard_normal((4000, 2))
    assert_rel(d.apply(x), O.spmm_csr(o, x), spmv_bound(o, x))
    for tpr_target, avg in [(2, 12), (8, 50), (32, 300)]:
        o2 = random_csr(rng, 300, 2000, rng.integers(max(avg - 5, 0), avg + 5, 300))
        d2 = to_dev(ctx, o2)
        assert d2.plan()["threads_per_row"] == tpr_target
        x2 = rng.standard_normal((2000, 1))
        assert_rel(d2.apply(x2), O.spmm_csr(o2, x2), spmv_bound(o2, x2))
    empty = to_dev(ctx, random_csr(rng, 10, 10, [0] * 10))
    assert np.array_equal(empty.apply(np.ones(10)), np.zeros((10, 1)))


def test_residual_add_and_fused_smoother(ctx, F, variant):
    o = O.gen_g7(20, 18, 16)
    d = to_dev(ctx, o)
    rng = np.random.default_rng(3)
    n = o.nrows
    for k in (1, 2):
        x, b = rng.standard_normal((n, k)), rng.standard_normal((n, k))
        X, B, R = F.DeviceMat.from_host(ctx, x), F.DeviceMat.from_host(ctx, b), F.DeviceMat(ctx, n, k)
        d.residual_dev(R, B, X)
        assert np.array_equal(R.to_host(), b - O.spmm_csr(o, x))          # multigrid.rs:341-342
        d.apply_add_dev(B, X)
        assert np.array_equal(B.to_host(), b + O.spmm_csr(o, x))          # multigrid.rs:349-350
        for kind, diag in [("l1", O.new_l1(o)), ("l2", O.new_l2(o)), ("jac", O.new_jacobi(o, 0.66))]:
            sm = {"l1": F.new_l1, "l2": F.new_l2}.get(kind, lambda m: F.new_jacobi(m, 0.66))(d)
            assert np.array_equal(sm.column_vector(), diag)               # smoothers.rs:43-86
            for iters in (1, 2, 3):
                X2, B2 = F.DeviceMat.from_host(ctx, x), F.DeviceMat.from_host(ctx, b)
                F.smooth(X2, B2, d, sm, iters)
                assert np.array_equal(X2.to_host(), O.smooth_diag(o, diag, x, b, iters))  # multigrid.rs:407-424


def test_fused_smoother_27pt_within_tolerance(ctx, F, variant):
    o = O.gen_g27(10, 9, 8)
    d = to_dev(ctx, o)
    rng = np.random.default_rng(4)
    x, b = rng.standard_normal((o.nrows, 1)), rng.standard_normal((o.nrows, 1))
    diag = O.new_l1(o)
    X, B = F.DeviceMat.from_host(ctx, x), F.DeviceMat.from_host(ctx, b)
    F.smooth(X, B, d, F.new_l1(d), 2)
    want = O.smooth_diag(o, diag, x, b, 2)
    assert np.max(np.abs(X.to_host() - want)) <= 1e-12 * np.max(np.abs(want))


def test_missing_diagonal_is_an_error(ctx, F):
    o = O.Csr.from_triplets(3, 3, [0, 1, 2], [0, 2, 2], [1.0, 1.0, 1.0])
    with pytest.raises(F.FamgError) as e:
        F.new_jacobi(to_dev(ctx, o), 0.66)  # mat.get(i,i).unwrap() panics, smoothers.rs:83
    assert e.value.status == F._ffi.ERR_NUMERIC


def test_stationary_iteration_literal(ctx, F, variant):
    o = O.gen_g7(9, 8, 7)
    d = to_dev(ctx, o)
    nn = np.random.default_rng(5).standard_normal((o.nrows, 3))
    got = F.StationaryIteration(d, F.new_l1(d), 3).apply(nn)
    assert np.array_equal(got, O.stationary_iteration(o, O.new_l1(o), 3, nn))  # smoothers.rs:146-159


def test_stationary_transpose_and_error_propagator(ctx, F):
    o = O.gen_g7(9, 8, 7)
    d = to_dev(ctx, o)
    rng = np.random.default_rng(21)
    x = rng.standard_normal((o.nrows, 3))
    l1 = O.new_l1(o)
    got = F.StationaryIteration(d, F.new_l1(d), 3).transpose_apply(x)
    assert np.array_equal(got, O.stationary_iteration_transpose(o, l1, 3, x))       # smoothers.rs:179-197
    e = F.ErrorPropogator(F.SparseMatOp(d), F.new_l1(d)).apply(x)
    assert np.array_equal(e, O.error_propagator(o, lambda r: l1.reshape(-1, 1) * r, x))  # adaptivity.rs:191-198


def test_degenerate_shapes(ctx, F):
    """Empty operators, zero right-hand sides, a single huge row."""
    empty = F.SparseRowMat.from_csr(ctx, 0, 0, [0], [], [])
    assert empty.apply(np.zeros((0, 2))).shape == (0, 2) and empty.transpose().shape == (0, 0)
    tall = F.SparseRowMat.from_csr(ctx, 3, 0, [0, 0, 0, 0], [], [])
    assert np.array_equal(tall.apply(np.zeros((0, 1))), np.zeros((3, 1)))
    assert (tall.transpose() @ tall).nnz == 0 and (tall.transpose() @ tall).shape == (0, 0)
    rng = np.random.default_rng(22)
    one = random_csr(rng, 1, 50000, [30000])
    dv = to_dev(ctx, one)
    x = rng.standard_normal((50000, 1))
    assert_rel(dv.apply(x), O.spmm_csr(one, x), spmv_bound(one, x))
    t = dv.transpose()
    ot = O.transpose(one)
    assert same_pattern(t, ot) and np.array_equal(t.to_host()[2], ot.val)
    c = dv @ t                                          # 1 x 1 product with 30000 terms, one output entry
    oc = O.spgemm(one, ot)
    assert c.nnz == 1 and np.array_equal(c.to_host()[2], oc.val)


def test_block_smoother(ctx, F):
    o = O.gen_g7(8, 6, 4)
    part, _ = F.geometric_partition((8, 6, 4))
    bs = F.BlockSmoother.new(F.SparseMatOp(to_dev(ctx, o)), part)
    r = np.random.default_rng(6).standard_normal((o.nrows, 2))
    want = O.block_smoother_apply(o, part.agg_ptr, part.agg_nodes, r)
    assert np.max(np.abs(bs.apply(r) - want)) <= 1e-12 * np.max(np.abs(want))


def test_block_smoother_device_build_equals_host_build(ctx, F):
    """The blocks are assembled, factorised and inverted on the device (one thread per aggregate) with the statements of the
    host routine in the same order: same bits.  Odd grid sizes give aggregates of 8, 4, 2 and 1 nodes."""
    dims = (10, 7, 5)
    o = O.gen_g7(*dims)
    part, _ = F.geometric_partition(dims)
    d = to_dev(ctx, o)
    r = np.random.default_rng(16).standard_normal((o.nrows, 3))
    dev = F.BlockSmoother.new(F.SparseMatOp(d), part).apply(r)
    ctx.set_option("block_build_host", 1)
    try:
        host = F.BlockSmoother.new(F.SparseMatOp(d), part).apply(r)
    finally:
        ctx.set_option("block_build_host", 0)
    assert np.array_equal(dev, host)
    want = O.block_smoother_apply(o, part.agg_ptr, part.agg_nodes, r)
    assert np.max(np.abs(dev - want)) <= 1e-12 * np.max(np.abs(want))
    # a block that is not positive definite is an error on both paths (block_smoothers.rs: the Cholesky factorisation fails)
    bad = to_dev(ctx, O.Csr.from_triplets(2, 2, [0, 0, 1, 1], [0, 1, 0, 1], [1.0, 2.0, 2.0, 1.0]))
    one = F.Partition.from_node_to_agg(np.zeros(2, dtype=np.int64))
    for host_build in (0, 1):
        ctx.set_option("block_build_host", host_build)
        try:
            with pytest.raises(F.FamgError):
                F.BlockSmoother.new(F.SparseMatOp(bad), one)
        finally:
            ctx.set_option("block_build_host", 0)


# ------------------------------------------------------------------ SpGEMM / transpose / RAP
def _check_product(ctx, F, oa, ob):
    c = to_dev(ctx, oa) @ to_dev(ctx, ob)
    oc = O.spgemm(oa, ob)
    assert same_pattern(c, oc), "structural (unpruned, sorted) pattern differs"
    assert np.array_equal(c.to_host()[2], oc.val), "values differ (ascending-k accumulation order)"


def test_spgemm_bit_exact_all_size_classes(ctx, F):
    rng = np.random.default_rng(7)
    # class 0 (ub <= 128), class 1 (<= 1024), class 2 (<= 8192, one CTA per row), class 3 (global tables)
    _check_product(ctx, F, random_csr(rng, 300, 200, rng.integers(0, 6, 300)), random_csr(rng, 200, 250, rng.integers(0, 8, 200)))
    _check_product(ctx, F, random_csr(rng, 120, 300, rng.integers(10, 30, 120)), random_csr(rng, 300, 400, rng.integers(5, 30, 300)))
    _check_product(ctx, F, random_csr(rng, 40, 400, rng.integers(60, 90, 40)), random_csr(rng, 400, 3000, rng.integers(40, 90, 400)))
    a = random_csr(rng, 12, 500, rng.integers(150, 200, 12))
    _check_product(ctx, F, a, random_csr(rng, 500, 6000, rng.integers(60, 110, 500)))
    # staged CTA-per-row accumulation: B rows longer than one staging batch (2048 entries: split
    # segments, whole-CTA accumulation), and thousands of short B rows per output row (64-segment
    # batches, single-warp accumulation)
    _check_product(ctx, F, random_csr(rng, 6, 300, rng.integers(100, 200, 6)), random_csr(rng, 300, 6000, rng.integers(1800, 5000, 300)))
    _check_product(ctx, F, random_csr(rng, 9, 4000, rng.integers(1000, 2500, 9)), random_csr(rng, 4000, 3000, rng.integers(0, 14, 4000)))
    # the same two shapes with B wider than the dense-accumulator limit (8192 columns): hash table,
    # sort and slot search, staged accumulation
    _check_product(ctx, F, random_csr(rng, 5, 300, rng.integers(100, 200, 5)), random_csr(rng, 300, 12000, rng.integers(300, 3000, 300)))
    _check_product(ctx, F, random_csr(rng, 7, 4000, rng.integers(1000, 2500, 7)), random_csr(rng, 4000, 12000, rng.integers(0, 14, 4000)))
    # explicit zeros are kept (cancellation must not prune)
    z = O.Csr.from_triplets(2, 2, [0, 0, 1], [0, 1, 1], [1.0, -1.0, 1.0])
    w = O.Csr.from_triplets(2, 1, [0, 1], [0, 0], [1.0, 1.0])
    c = to_dev(ctx, z) @ to_dev(ctx, w)
    assert c.nnz == 2 and c.to_host()[2].tolist() == [0.0, 1.0]


def _banded_csr(rng, nrows, ncols, lens, starts, width):
    """Random CSR whose row i draws lens[i] sorted unique columns from [starts[i], starts[i] + width)."""
    rp = np.zeros(nrows + 1, dtype=np.int64)
    cols, vals = [], []
    for i in range(nrows):
        lo = int(min(starts[i], ncols - 1)); hi = int(min(lo + width, ncols))
        k = int(min(lens[i], hi - lo))
        cols.append(lo + np.sort(rng.choice(hi - lo, size=k, replace=False)))
        vals.append(rng.standard_normal(k))
        rp[i + 1] = rp[i] + k
    return O.Csr.from_arrays(nrows, ncols, rp, np.concatenate(cols), np.concatenate(vals))


def _window_case(seed=17, blens=(60, 140)):
    """A (220 x 3000) times a wide banded B (3000 x 25500, blens entries per row): thousands of products per row, all candidate columns of a row inside a
    window of <= 6300 columns that moves with the row (the window class of spgemm.cu), except every 7th row of A, which is
    spread over all of B (no window: table classes) -- both kinds in one product."""
    rng = np.random.default_rng(seed)
    m, inner = 220, 3000
    starts = np.minimum(np.arange(m) * 10, inner - 600)
    a = _banded_csr(rng, m, inner, rng.integers(130, 170, m), starts, 600)
    wide = random_csr(rng, m, inner, rng.integers(130, 170, m))
    rp, col, val = [0], [], []
    for i in range(m):
        src = wide if i % 7 == 3 else a
        col.append(src.col[src.row_ptr[i]:src.row_ptr[i + 1]]); val.append(src.val[src.row_ptr[i]:src.row_ptr[i + 1]])
        rp.append(rp[-1] + len(col[-1]))
    a = O.Csr.from_arrays(m, inner, np.array(rp, dtype=np.int64), np.concatenate(col), np.concatenate(val))
    lens = rng.integers(blens[0], blens[1], inner)
    b = _banded_csr(rng, inner, 25500, lens, np.arange(inner) * 8, 1500)
    return a, b


def test_spgemm_window_rows(ctx, F):
    _check_product(ctx, F, *_window_case())
    rng = np.random.default_rng(18)
    # rows of B longer than several staging chunks (pieces of one B_k run one after the other), window = all of B
    a = random_csr(rng, 4, 64, rng.integers(20, 60, 4))
    _check_product(ctx, F, a, random_csr(rng, 64, 8000, rng.integers(3000, 7000, 64)))
    # 1 000 - 2 000 rows of B per output row (dozens of descriptor groups), some of them empty
    a = random_csr(rng, 6, 4000, rng.integers(1000, 2000, 6))
    _check_product(ctx, F, a, _banded_csr(rng, 4000, 9000, rng.integers(0, 200, 4000), rng.integers(0, 8000, 4000), 1000))


def _run_product_check(env_extra, body):
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import numpy as np, oracle as O, faer_amg_b200 as F\n"
            "from util import to_dev, same_pattern, random_csr\n"
            "from test_gpu_parity import _window_case, _banded_csr\n"
            "ctx = F.Context.default(0)\n"
            "def check(a, b):\n"
            "    c = to_dev(ctx, a) @ to_dev(ctx, b); oc = O.spgemm(a, b)\n"
            "    assert same_pattern(c, oc) and np.array_equal(c.to_host()[2], oc.val)\n"
            "%s\n"
            "print('product ok')\n") % (root, os.path.join(root, "tests"), body)
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env_extra), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "product ok" in out.stdout, out.stdout + out.stderr


def test_spgemm_window_rows_deferred_to_fill_pass():
    """With a scratch buffer of 2 000 entries most window rows do not fit: the fill pass recomputes them in place."""
    _run_product_check({"FAMG_SG_WIN_CAP": "2000"}, "check(*_window_case())")


def test_spgemm_window_rows_short_segments():
    """The window kernel itself on rows it is normally not given (B rows of 0 - 60 and of 1 - 3 entries)."""
    _run_product_check({"FAMG_SG_WIN_MIN_SEG": "0"},
                       "check(*_window_case(19, (0, 60)))\n"
                       "rng = np.random.default_rng(20)\n"
                       "a = random_csr(rng, 5, 5000, rng.integers(1800, 4000, 5))\n"
                       "check(a, _banded_csr(rng, 5000, 700, rng.integers(1, 4, 5000), rng.integers(0, 600, 5000), 100))")


def test_transpose_bit_exact(ctx, F):
    rng = np.random.default_rng(8)
    for o in [random_csr(rng, 300, 120, rng.integers(0, 9, 300)), O.gen_g27(6), random_csr(rng, 5, 2000, [900, 0, 1500, 3, 40])]:
        t = to_dev(ctx, o).transpose()
        ot = O.transpose(o)
        assert same_pattern(t, ot) and np.array_equal(t.to_host()[2], ot.val)


@pytest.mark.parametrize("gen,dims", [(O.gen_g7, (12, 10, 8)), (O.gen_g27, (8, 8, 6))])
def test_smoothed_aggregation_galerkin(ctx, F, gen, dims):
    """interpolation/mod.rs:730-836 end to end: tentative P, smoothing, transpose, R(AP)."""
    o = gen(*dims)
    n = o.nrows
    rng = np.random.default_rng(9)
    nn = 1.0 + 0.05 * rng.standard_normal((n, 1))
    part, _ = F.geometric_partition(dims)
    for steps in (0, 1, 2):
        cnn, r, p, ac, _ = F.smoothed_aggregation(to_dev(ctx, o), part, 1, nn, 1, steps)
        g = O.smoothed_aggregation(o, part.agg_ptr, part.agg_nodes, nn, 1, steps)
        for dev, orc in [(p, g.interpolation), (r, g.restriction), (ac, g.coarse_mat)]:
            assert same_pattern(dev, orc)
            assert O.mats_are_equal(to_oracle(dev), orc)          # utils.rs:32-58 (1e-12 abs & rel)
            assert np.array_equal(dev.to_host()[2], orc.val)      # and in fact identical bits
        assert np.array_equal(cnn, g.coarse_nn)


def test_tentative_p_multi_candidate(ctx, F):
    rng = np.random.default_rng(10)
    dims = (6, 4, 4)
    n = int(np.prod(dims))
    nn = rng.standard_normal((n, 3))
    part, _ = F.geometric_partition(dims)
    p, cnn = F.tentative_prolongator(ctx, n, part, nn, 2)
    op, ocnn = O.tentative_p(n, nn, part.agg_ptr, part.agg_nodes, cand=2)
    assert same_pattern(p, op) and np.array_equal(p.to_host()[2], op.val) and np.array_equal(cnn, ocnn)
    small = F.Partition.from_node_to_agg(np.arange(n))  # singleton aggregates cannot carry 2 candidates
    with pytest.raises(F.FamgError):
        F.tentative_prolongator(ctx, n, small, nn, 2)     # interpolation/mod.rs:757-762


def test_smooth_interpolation_errors(ctx, F):
    o = O.Csr.from_triplets(2, 2, [0, 0, 1], [0, 1, 0], [1.0, 1.0, 1.0])  # row 1 has no diagonal
    p = O.Csr.from_triplets(2, 1, [0, 1], [0, 0], [1.0, 1.0])
    with pytest.raises(F.FamgError):
        F.smooth_interpolation(to_dev(ctx, o), to_dev(ctx, p))            # interpolation/mod.rs:933-934


# ------------------------------------------------------------------ hierarchy / cycle / solve
def _build_both(ctx, F, gen, dev_gen, dims, smoother, coarsest_dim=1000, mu=1, nu=1):
    o = gen(*dims)
    n = o.nrows
    nn = np.full((n, 1), 1.0 / np.sqrt(n))
    oh = O.build_hierarchy(o, nn, dims, coarsest_dim=coarsest_dim)
    omg = O.multigrid_from_hierarchy(oh, smoother, mu=mu, nu=nu)
    a = dev_gen(ctx, *dims)
    cfg = F.HierarchyConfig(coarsest_dim, F.AggregationConfig(1, 1, F.GeometricPartitioner(dims)))
    h = cfg.build(F.SparseMatOp(a), nn)
    mg = F.MultigridConfig(mu=mu, smoothing_steps=nu, smoother=smoother).build(h)
    return o, oh, omg, a, h, mg


def test_hierarchy_structure_bit_exact(ctx, F):
    o, oh, omg, a, h, mg = _build_both(ctx, F, O.gen_g7, F.gallery.poisson7, (16, 16, 16), "l1", coarsest_dim=100)
    assert h.levels() == oh.levels == 3
    for lvl in range(h.levels()):
        assert same_pattern(h.get_mat_ref(lvl), oh.operators[lvl])          # bit-exact sparsity on every level
        assert O.mats_are_equal(to_oracle(h.get_mat_ref(lvl)), oh.operators[lvl])   # utils.rs:32-58, 1e-12
        scale = np.max(np.abs(oh.near_nulls[lvl]))
        assert np.max(np.abs(h.get_near_null(lvl) - oh.near_nulls[lvl])) <= 1e-12 * scale
    # level 1 is built from bit-identical inputs and SpGEMM keeps the reference's summation order
    assert np.array_equal(h.get_mat_ref(1).to_host()[2], oh.operators[1].val)
    for lvl in range(h.levels() - 1):
        assert same_pattern(h.get_interpolation(lvl), oh.interpolations[lvl])
        assert same_pattern(h.get_restriction(lvl), oh.restrictions[lvl])
        assert O.mats_are_equal(to_oracle(h.get_interpolation(lvl)), oh.interpolations[lvl])
    assert h.op_complexity() == oh.op_complexity() and h.grid_complexity() == oh.grid_complexity()


@pytest.mark.parametrize("mu,nu,smoother", [(1, 1, "l1"), (1, 2, "jacobi"), (2, 1, "l2"), (2, 2, "l1")])
def test_multigrid_apply_matches_oracle(ctx, F, variant, mu, nu, smoother):
    o, oh, omg, a, h, mg = _build_both(ctx, F, O.gen_g7, F.gallery.poisson7, (16, 12, 8), smoother, coarsest_dim=60, mu=mu, nu=nu)
    assert mg.levels() == oh.levels >= 3
    rhs = np.random.default_rng(11).standard_normal((o.nrows, 3))
    got, want = mg.apply(rhs), omg.apply(rhs)
    # identical operation order on every level except the coarsest solve (explicit inverse vs
    # triangular solves): relative to the result's scale the difference is ~1e-15
    assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))
    # device-resident (CUDA-graph) path gives the same bits as the staged path, call after call
    R, Z = F.DeviceMat.from_host(ctx, rhs[:, :1]), F.DeviceMat(ctx, o.nrows, 1)
    mg.apply_dev(Z, R); first = Z.to_host()
    mg.apply_dev(Z, R)
    assert np.array_equal(first, Z.to_host()) and np.array_equal(first, got[:, :1])


def test_multigrid_symmetry(ctx, F):
    """symmetry_test (multigrid.rs:520-580) on the device operator."""
    o, oh, omg, a, h, mg = _build_both(ctx, F, O.gen_g27, F.gallery.diffusion27, (8, 8, 8), "l1", coarsest_dim=100)
    rng = np.random.default_rng(12)
    u, v = rng.standard_normal((o.nrows, 5)), rng.standard_normal((o.nrows, 5))
    utbv, vtbu = u.T @ mg.apply(v), v.T @ mg.apply(u)
    assert np.max(np.abs(utbv - vtbu.T)) <= 1e-12 * np.max(np.abs(utbv))


def test_block_smoother_multigrid(ctx, F):
    """MultigridConfig::default() wiring: BlockSmoother on every level but the coarsest."""
    dims = (8, 8, 8)
    o, oh, omg, a, h, _ = _build_both(ctx, F, O.gen_g7, F.gallery.poisson7, dims, "l1", coarsest_dim=100)
    mg = F.MultigridConfig(smoother="block").build(h)
    ob = O.Multigrid()
    for lvl, m in enumerate(oh.operators):
        if lvl == oh.levels - 1:
            sm = "cholesky"
        else:
            sm = ("block", oh.partitions[lvl][0], oh.partitions[lvl][1])
        ob.add_level(m, sm, *( (oh.restrictions[lvl - 1], oh.interpolations[lvl - 1]) if lvl else (None, None)))
    rhs = np.random.default_rng(13).standard_normal((o.nrows, 2))
    got, want = mg.apply(rhs), ob.apply(rhs)
    assert np.max(np.abs(got - want)) <= 1e-11 * np.max(np.abs(want))


def test_simple_geometric_example_on_gpu(ctx, F):
    """examples/simple_geometric.rs (BASELINE config #1) through the GPU path vs the golden table."""
    for row in GOLD["simple_geometric"][:6]:
        ne = row["dofs"] + 1
        refinement = int(np.log2(ne // 10))
        a = F.gallery.poisson1d(ctx, ne)
        mg = F.Multigrid(a, F.new_jacobi(a, 0.66))
        for level in range(1, refinement + 1):
            ce = 10 * 2 ** (refinement - level)
            m = F.gallery.poisson1d(ctx, ce)
            sm = F.SparseCholeskySolve.new(m) if level == refinement else F.new_jacobi(m, 0.66)
            mg.add_level(m, sm, F.gallery.restriction1d(ctx, ce - 1), F.gallery.interpolation1d(ctx, ce - 1))
        b = np.ones(ne - 1)
        params = F.CgParams(abs_tolerance=EPS, rel_tolerance=1e-8, max_iters=6000)
        x = np.zeros(ne - 1)
        i1 = F.conjugate_gradient(x, F.new_jacobi(a, 0.66), a, b, params).iter_count
        x[:] = 0
        i2 = F.conjugate_gradient(x, mg, a, b, params).iter_count
        x[:] = 0
        i3 = F.stationary_solver(x, b, a, mg, 6000, 1e-8)
        assert abs(i1 - row["pcg_jacobi"]) <= 1 and abs(i2 - row["pcg_mg"]) <= 1 and abs(i3 - row["stat_mg"]) <= 1


@pytest.mark.parametrize("case", ["g7_32_l1", "g7_48x32x16_jacobi", "g27_24_l1"])
def test_pcg_amg_iteration_counts(ctx, F, variant, case):
    g = GOLD["amg"][case]
    dims = tuple(g["dims"])
    gen = (O.gen_g27, F.gallery.diffusion27) if case.startswith("g27") else (O.gen_g7, F.gallery.poisson7)
    o, oh, omg, a, h, mg = _build_both(ctx, F, gen[0], gen[1], dims, g["smoother"])
    assert [m.mat_ref().nrows for m in h.operators()] == g["level_rows"]
    assert [m.mat_ref().nnz for m in h.operators()] == g["level_nnz"]
    b = np.ones(o.nrows)
    for tol in ("1e-08", "1e-12"):
        x = np.zeros(o.nrows)
        info = F.conjugate_gradient(x, mg, a, b, F.CgParams(0.0, float(tol), 1000))
        assert abs(info.iter_count - g["iters"][tol]["iters"]) <= 1
        xo, oinfo = O.pcg(o, b, omg, rel_tol=float(tol))
        assert abs(info.iter_count - oinfo.iters) <= 1
        assert np.linalg.norm(x - xo) <= 10 * float(tol) * np.linalg.norm(xo)
        assert np.linalg.norm(b - O.spmm_csr(o, x).ravel()) < float(tol) * np.linalg.norm(b) * 1.01


def test_pcg_error_paths(ctx, F):
    a = F.gallery.poisson7(ctx, 12)
    b = np.ones(a.nrows)
    x = np.zeros(a.nrows)
    with pytest.raises(F.CgError) as e:
        F.conjugate_gradient(x, None, a, b, F.CgParams(0.0, 1e-14, 3))
    assert e.value.kind == "NoConvergence" and e.value.iter_count == 3
    assert F.test_solver(a, None, None, b, 3, 1e-14)[0] == 1000                  # report_cg, utils.rs:647-656
    info = F.conjugate_gradient(x, None, a, np.zeros(a.nrows), F.CgParams())
    assert info.iter_count == 0 and not x.any()
    x0 = np.linalg.solve(a.to_scipy().toarray()[:1, :1], [1.0])  # noqa: F841 (shape smoke)
    xg = np.random.default_rng(14).standard_normal(a.nrows)
    info = F.conjugate_gradient(xg, F.new_jacobi(a, 1.0), a, b, F.CgParams(0.0, 1e-10, 500, initial_guess_zero=False))
    assert np.linalg.norm(b - a.to_scipy() @ xg) < 1.01e-10 * np.linalg.norm(b)


# ------------------------------------------------------------------ full-size, size-independent properties
def test_full_size_properties_128(ctx, F):
    """BASELINE config #2 size (128^3): properties that need no CPU oracle at that size."""
    n1 = 128
    a = F.gallery.poisson7(ctx, n1)
    n = n1 ** 3
    assert a.nnz == 7 * n - 6 * n1 * n1
    ones = F.DeviceMat.from_host(ctx, np.ones(n))
    y = F.DeviceMat(ctx, n, 1)
    a.apply_dev(y, ones)
    yh = y.to_host().reshape(n1, n1, n1)
    assert np.all(yh[1:-1, 1:-1, 1:-1] == 0.0) and yh.sum() == 6.0 * n1 * n1       # A 1 lives on the boundary
    rng = np.random.default_rng(15)
    u, v = rng.standard_normal(n), rng.standard_normal(n)
    au, av = a.apply(u).ravel(), a.apply(v).ravel()
    assert abs(v @ au - u @ av) <= 1e-12 * abs(v @ au)                             # symmetry
    assert np.max(np.abs(a.apply(u + v).ravel() - au - av)) <= 1e-12 * np.max(np.abs(au))   # linearity
    x_int = np.zeros((n1, n1, n1)); x_int[1:-1, 1:-1, 1:-1] = rng.standard_normal((n1 - 2,) * 3)
    lap = a.apply(x_int.ravel()).reshape(n1, n1, n1)[2:-2, 2:-2, 2:-2]                  # 7-point stencil, checked with numpy
    ref = 6 * x_int[2:-2, 2:-2, 2:-2] - x_int[1:-3, 2:-2, 2:-2] - x_int[3:-1, 2:-2, 2:-2] - x_int[2:-2, 1:-3, 2:-2] \
        - x_int[2:-2, 3:-1, 2:-2] - x_int[2:-2, 2:-2, 1:-3] - x_int[2:-2, 2:-2, 3:-1]
    assert np.max(np.abs(lap - ref)) <= 1e-12 * np.max(np.abs(ref))
    nn = np.full((n, 1), 1.0 / np.sqrt(n))
    h = F.HierarchyConfig(1000, F.AggregationConfig(1, 1, F.GeometricPartitioner((n1,) * 3))).build(F.SparseMatOp(a), nn)
    assert [op.mat_ref().nrows for op in h.operators()] == [n, n // 8, n // 64, n // 512, n // 4096]
    for lvl in range(h.levels() - 1):
        ac = h.get_mat_ref(lvl + 1).to_scipy()
        assert abs(ac - ac.T).max() <= 1e-13 * abs(ac).max()                       # RAP symmetric
        p = h.get_interpolation(lvl)
        r = h.get_restriction(lvl)
        assert (r.nrows, r.ncols) == (p.ncols, p.nrows) and r.nnz == p.nnz
    key = "g7_128_l1"
    mg = F.MultigridConfig(smoother="l1").build(h)
    x = np.zeros(n)
    info = F.conjugate_gradient(x, mg, a, np.ones(n), F.CgParams(0.0, 1e-8, 1000))
    assert info.rel_residual < 1e-8
    if key in GOLD["amg"]:
        g = GOLD["amg"][key]
        assert [op.mat_ref().nnz for op in h.operators()] == g["level_nnz"]
        assert abs(info.iter_count - g["iters"]["1e-08"]["iters"]) <= 1
        assert abs(np.linalg.norm(x) - g["iters"]["1e-08"]["x_norm"]) <= 1e-6 * g["iters"]["1e-08"]["x_norm"]
    else:
        assert 10 <= info.iter_count <= 25


def test_cpp_mirror_on_gpu():
    """include/famg.hpp end to end (Galerkin product, two-level multigrid, PCG) through the C ABI from C++."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "cpp", "mirror_smoke")
    if not os.path.exists(exe):
        pytest.skip("run __graft_entry__.build() first")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "mirror ok" in out.stdout and "distributed ok" in out.stdout, out.stdout + out.stderr
