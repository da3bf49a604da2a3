"""Distributed hierarchy construction (csrc/dist_setup.cu) on ONE GPU: a virtual communicator hosts all
ranks in this process, every exchange is a device copy, and the row slabs of P, R and A_c that the ranks
build must be bit-identical to the rows of the undistributed build (same SpGEMM kernels, same ascending
inner-index accumulation; the coarse near-null normalisation is chained through the ranks in order).
Reference: Hierarchy::coarsen (hierarchy.rs:190-248), smoothed_aggregation (interpolation/mod.rs:730-836)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _global_hierarchy(ctx, dims, stencil, coarsest):
    import faer_amg_b200 as F
    a = F.gallery.poisson7(ctx, *dims) if stencil == 7 else F.gallery.diffusion27(ctx, *dims)
    n = a.nrows
    nn = np.full((n, 1), 1.0 / np.sqrt(n))
    h = F.HierarchyConfig(coarsest, F.AggregationConfig(1, 1, F.GeometricPartitioner(dims))).build(F.SparseMatOp(a), nn)
    return a, nn, h


def _dist_hierarchy(ctx, dims, stencil, nranks, coarsest, rep):
    import faer_amg_b200 as F
    from faer_amg_b200.distributed import Comm, DistGeometricPartitioner, DistHierarchy, DistMat, fine_plane_splits
    comm = Comm.sim(ctx, nranks)
    rs = fine_plane_splits(dims, nranks)
    plane = dims[0] * dims[1]
    n = plane * dims[2]
    gen = F.gallery.poisson7_slab if stencil == 7 else F.gallery.diffusion27_slab
    slabs = [gen(ctx, *dims, int(rs[r]) // plane, int(rs[r + 1]) // plane) for r in range(nranks)]
    a0 = DistMat.from_slabs(comm, slabs, n)
    nn = [np.full(int(rs[r + 1] - rs[r]), 1.0 / np.sqrt(n)) for r in range(nranks)]
    return comm, rs, DistHierarchy(comm, a0, nn, DistGeometricPartitioner(dims), coarsest_dim=coarsest, replicate_below=rep)


def _rows(m, r0, r1):
    rp, ci, v = m.to_host()
    rp = rp.astype(np.int64)
    return rp[r0:r1 + 1] - rp[r0], ci[rp[r0]:rp[r1]].astype(np.int64), v[rp[r0]:rp[r1]]


def _same(slab, glob, r0, r1, what):
    rp, ci, v = slab.to_host()
    grp, gci, gv = _rows(glob, r0, r1)
    assert slab.nrows == r1 - r0 and slab.ncols == glob.ncols, what
    assert np.array_equal(rp.astype(np.int64), grp), f"{what}: row pointers differ"
    assert np.array_equal(ci.astype(np.int64), gci), f"{what}: pattern differs"
    assert np.array_equal(v, gv), f"{what}: values differ (max {np.max(np.abs(v - gv)):.3e})"


@pytest.mark.parametrize("nranks,dims,stencil,coarsest,rep", [
    (2, (16, 16, 16), 7, 40, 200),      # one distributed level
    (4, (16, 16, 32), 7, 20, 64),       # two distributed levels, four ranks
    (3, (12, 10, 24), 27, 30, 100),     # 27-point operator, uneven slabs (24 planes over 3 ranks, granule 8)
    (8, (8, 8, 64), 7, 10, 16),         # eight thin slabs
])
def test_dist_hierarchy_bit_identical_to_single_gpu(ctx, nranks, dims, stencil, coarsest, rep):
    a, nn, h = _global_hierarchy(ctx, dims, stencil, coarsest)
    comm, rs, dh = _dist_hierarchy(ctx, dims, stencil, nranks, coarsest, rep)
    assert dh.levels() == h.levels()
    nd = dh.n_dist
    assert nd >= 1
    for l in range(nd):
        rsl = dh.A[l].row_split()
        csl = dh.P[l].info()[3]
        for r in range(nranks):
            _same(dh.A[l].local(r, True), h.get_mat_ref(l), int(rsl[r]), int(rsl[r + 1]), f"A level {l} rank {r}")
            _same(dh.P[l].local(r, True), h.get_interpolation(l), int(rsl[r]), int(rsl[r + 1]), f"P level {l} rank {r}")
            _same(dh.R[l].local(r, True), h.get_restriction(l), int(csl[r]), int(csl[r + 1]), f"R level {l} rank {r}")
        if l > 0:
            got = dh.near_null_host(l)
            assert np.array_equal(got, h.get_near_null(l).ravel()), f"near-null level {l}"
    # replicated tail == remaining levels of the undistributed hierarchy
    for t in range(dh.tail.levels()):
        g, w = dh.tail.get_mat_ref(t), h.get_mat_ref(nd + t)
        _same(g, w, 0, w.nrows, f"tail A level {nd + t}")
        assert np.array_equal(dh.tail.get_near_null(t).ravel(), h.get_near_null(nd + t).ravel()), f"tail near-null {nd + t}"
        if t > 0:
            _same(dh.tail.get_interpolation(t - 1), h.get_interpolation(nd + t - 1), 0, h.get_interpolation(nd + t - 1).nrows, "tail P")


def test_dist_matrix_gather_and_plan(ctx):
    """Halo plan bookkeeping: renumbered slabs map back to the global operator; gather reproduces it."""
    import faer_amg_b200 as F
    from faer_amg_b200.distributed import Comm, DistMat, fine_plane_splits
    dims, nranks = (6, 5, 12), 3
    a = F.gallery.diffusion27(ctx, *dims)
    comm = Comm.sim(ctx, nranks)
    rs = fine_plane_splits(dims, nranks)
    plane = dims[0] * dims[1]
    slabs = [F.gallery.diffusion27_slab(ctx, *dims, int(rs[r]) // plane, int(rs[r + 1]) // plane) for r in range(nranks)]
    m = DistMat.from_slabs(comm, slabs, a.nrows).finalize(False)
    nr, nc, rsm, csm = m.info()
    assert (nr, nc) == a.shape and np.array_equal(rsm, rs) and np.array_equal(csm, rs)
    for r in range(nranks):
        loc = m.local(r, False)
        assert loc.ncols <= a.ncols and loc.ncols >= int(rs[r + 1] - rs[r])  # [owned | ghost]
        _same(m.local(r, True), a, int(rs[r]), int(rs[r + 1]), f"rank {r}")
    _same(m.gather(), a, 0, a.nrows, "gather")


def test_virtual_communicator_cannot_run_the_cycle(ctx):
    import faer_amg_b200 as F
    from faer_amg_b200.distributed import Comm, DistMultigrid
    comm, rs, dh = _dist_hierarchy(ctx, (16, 16, 16), 7, 2, 40, 200)
    with pytest.raises(F.FamgError):
        DistMultigrid.from_hierarchy(comm, dh)


def test_tail_of_one_level(ctx):
    """The first replicated level is already below coarsest_dim: the tail is that single level (no extra coarsening)."""
    a, nn, h = _global_hierarchy(ctx, (24, 20, 16), 7, 200)
    comm, rs, dh = _dist_hierarchy(ctx, (24, 20, 16), 7, 2, 200, 100)
    assert dh.levels() == h.levels() == 3 and dh.tail.levels() == 1
    w = h.get_mat_ref(2)
    _same(dh.tail.get_mat_ref(0), w, 0, w.nrows, "tail")


@pytest.mark.parametrize("dims,block", [((7, 5, 9), (2, 2, 2)), ((16, 16, 16), (2, 2, 2)), ((9, 4, 6), (3, 2, 1)), ((1, 1, 5), (2, 2, 2))])
def test_device_aggregates_and_tentative_p_bit_identical_to_host(ctx, dims, block):
    """famg_partition_geometric_dev == famg_geometric_partition, famg_tentative_p_dev == famg_tentative_p (bit for bit),
    upload / download round trip."""
    import faer_amg_b200 as F
    from faer_amg_b200.interpolation import tentative_prolongator, tentative_prolongator_dev
    from faer_amg_b200.partitioners import DevicePartition, geometric_partition
    hp, hc = geometric_partition(dims, block)
    dp, dc = DevicePartition.geometric(ctx, dims, block)
    assert hc == dc and dp.naggs() == hp.naggs() and dp.nnodes() == hp.nnodes()
    got = dp.to_host()
    assert np.array_equal(got.agg_ptr, hp.agg_ptr) and np.array_equal(got.agg_nodes, hp.agg_nodes)
    up = DevicePartition.from_host(ctx, hp)
    up._host = None
    back = up.to_host()
    assert np.array_equal(back.agg_ptr, hp.agg_ptr) and np.array_equal(back.agg_nodes, hp.agg_nodes)
    n = hp.nnodes()
    rng = np.random.default_rng(3)
    nn = rng.standard_normal((n, 1))
    p_h, cnn_h = tentative_prolongator(ctx, n, hp, nn, 1, 1)
    p_d, cnn_d = tentative_prolongator_dev(dp, F.DeviceMat.from_host(ctx, nn))
    a, b = p_h.to_host(), p_d.to_host()
    assert p_h.shape == p_d.shape and all(np.array_equal(x, y) for x, y in zip(a, b))
    assert np.array_equal(cnn_h.ravel(), cnn_d.to_host().ravel())


def test_device_tentative_p_rejects_bad_partitions(ctx):
    import faer_amg_b200 as F
    from faer_amg_b200.interpolation import tentative_prolongator_dev
    from faer_amg_b200.partitioners import DevicePartition, Partition
    nn = F.DeviceMat.from_host(ctx, np.ones((4, 1)))
    twice = Partition(np.array([0, 2, 4]), np.array([0, 1, 1, 3]), 4, validate=False)   # node 1 twice, node 2 never
    with pytest.raises(F.FamgError):
        tentative_prolongator_dev(DevicePartition.from_host(ctx, twice), nn)
    empty = Partition(np.array([0, 0, 4]), np.array([0, 1, 2, 3]), 4, validate=False)   # interpolation/mod.rs:757-762
    with pytest.raises(F.FamgError):
        tentative_prolongator_dev(DevicePartition.from_host(ctx, empty), nn)


def test_hierarchy_host_and_device_aggregate_paths_agree(ctx):
    """Hierarchy::coarsen with the aggregates produced on the host (callable partitioner) and on the device."""
    import faer_amg_b200 as F
    dims = (12, 10, 8)
    a = F.gallery.diffusion27(ctx, *dims)
    nn = np.full((a.nrows, 1), 1.0 / np.sqrt(a.nrows))
    gp = F.GeometricPartitioner(dims)
    h_dev = F.HierarchyConfig(30, F.AggregationConfig(1, 1, gp)).build(F.SparseMatOp(a), nn)
    gp2 = F.GeometricPartitioner(dims)
    h_host = F.HierarchyConfig(30, F.AggregationConfig(1, 1, lambda level, op, v: gp2(level, op, v))).build(F.SparseMatOp(a), nn)
    assert h_dev.levels() == h_host.levels() >= 3
    for l in range(1, h_dev.levels()):
        _same(h_dev.get_mat_ref(l), h_host.get_mat_ref(l), 0, h_host.get_mat_ref(l).nrows, f"A level {l}")
        _same(h_dev.get_interpolation(l - 1), h_host.get_interpolation(l - 1), 0, h_host.get_interpolation(l - 1).nrows, f"P level {l - 1}")
        assert np.array_equal(h_dev.get_near_null(l), h_host.get_near_null(l))
        assert np.array_equal(np.asarray(h_dev.get_partition(l - 1).agg_nodes), np.asarray(h_host.get_partition(l - 1).agg_nodes))
