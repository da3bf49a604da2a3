"""CPU tests of the host partitioner: the C++ restatement in libfamg (host-only entry
points, no GPU needed) against the pure-Python oracle ``oracle/partitioner.py`` -- bit-exact
strength graphs and aggregates under the documented tie-breaking -- plus the structural properties
the reference's algorithm guarantees (``Partition::validate``, partitioners/mod.rs:144-154)."""
import numpy as np
import pytest

import oracle as O
from oracle import partitioner as OP
from faer_amg_b200.partitioners import Partition, PartitionerConfig, StrengthGraph


def _pattern(a: O.Csr):
    return np.asarray(a.row_ptr, dtype=np.int64), np.asarray(a.col, dtype=np.int64)


def _graph_lists(g: StrengthGraph):
    rp, ci, w = g.to_csr()
    return [[(int(ci[q]), float(w[q])) for q in range(rp[i], rp[i + 1])] for i in range(len(rp) - 1)]


def _near_null(kind, n, k, seed=0):
    if kind == "const":
        return np.full((n, k), 1.0 / np.sqrt(n))
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, k))


CASES = [
    ("g7-5x4x3-random", lambda: O.gen_g7(5, 4, 3), "rand", 2),
    ("g7-6-const", lambda: O.gen_g7(6), "const", 1),          # every distance ties (F9's worst case)
    ("g27-4-random", lambda: O.gen_g27(4), "rand", 3),
    ("g1-40-random", lambda: O.gen_g1(41), "rand", 1),
]


@pytest.mark.parametrize("name,gen,kind,k", CASES, ids=[c[0] for c in CASES])
def test_strength_graph_bit_exact(name, gen, kind, k):
    a = gen()
    rp, ci = _pattern(a)
    nn = _near_null(kind, a.nrows, k)
    w = 1.0 / (1.0 + np.arange(k))
    want = OP.new_ls_strength_graph(rp, ci, nn, w, 3)
    got = _graph_lists(StrengthGraph.new_ls_strength_graph((rp, ci), nn, w, 3))
    assert got == want  # ids and weights, bit for bit
    for i, nb in enumerate(got):
        assert nb and all(j != i for j, _ in nb) and all(0.0 <= x <= 1.0 for _, x in nb)
        assert [j for j, _ in nb] == sorted({j for j, _ in nb})


@pytest.mark.parametrize("name,gen,kind,k", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("cf", [2.0, 8.0])
def test_partition_bit_exact_and_valid(name, gen, kind, k, cf):
    a = gen()
    rp, ci = _pattern(a)
    nn = _near_null(kind, a.nrows, k, seed=1)
    w = np.ones(k)
    want, _ = OP.build_partition(rp, ci, nn, w, coarsening_factor=cf, max_improvement_iters=20)
    cfg = PartitionerConfig(coarsening_factor=cf, max_improvement_iters=20)
    part = cfg.build_from_strength(StrengthGraph.new_ls_strength_graph((rp, ci), nn, w, 3))
    assert np.array_equal(part.node_to_agg(), want)
    part.validate()
    assert part.nnodes() == a.nrows
    # greedy matching stops once nnodes / naggs >= cf, and each round at most halves the count
    assert part.nnodes() / part.naggs() >= min(cf, part.nnodes()) or part.naggs() == 1
    assert part.nnodes() / part.naggs() < 2.0 * cf + 1.0


def test_refinement_only_moves_with_positive_gain():
    # improve_partition never changes the number of aggregates (modularity.rs:449-453)
    a = O.gen_g7(6, 5, 4)
    rp, ci = _pattern(a)
    nn = _near_null("rand", a.nrows, 2, seed=3)
    g = StrengthGraph.new_ls_strength_graph((rp, ci), nn, np.ones(2), 3)
    p0 = PartitionerConfig(4.0, 1.0, 0).build_from_strength(g)
    p1 = PartitionerConfig(4.0, 1.0, 50).build_from_strength(g)
    assert p0.naggs() == p1.naggs()
    w0, _ = OP.build_partition(rp, ci, nn, np.ones(2), 4.0, 1.0, 0)
    assert np.array_equal(p0.node_to_agg(), w0)


def test_explicit_graph_round_trip_and_errors():
    from faer_amg_b200._ffi import FamgError

    rp = np.array([0, 1, 3, 4]); ci = np.array([1, 0, 2, 1]); w = np.array([1.0, 1.0, 0.5, 0.5])
    g = StrengthGraph.from_csr(rp, ci, w)
    assert g.dims() == (3, 4)
    r2, c2, w2 = g.to_csr()
    assert r2.tolist() == rp.tolist() and c2.tolist() == ci.tolist() and w2.tolist() == w.tolist()
    part = PartitionerConfig(2.0, 1.0, 10).build_from_strength(g)
    part.validate()
    # an isolated node has an empty neighbourhood: the reference panics "graph is disconnected" (mod.rs:373)
    with pytest.raises(FamgError):
        StrengthGraph.new_ls_strength_graph((np.array([0, 1, 2]), np.array([0, 1])), np.ones((2, 1)), [1.0], 3)
    with pytest.raises(ValueError):
        OP.new_ls_strength_graph(np.array([0, 1, 2]), np.array([0, 1]), np.ones((2, 1)), [1.0], 3)


def test_partition_scales_to_a_moderate_grid():
    # 24^3 = 13,824 nodes, constant near-null: exercises the OpenMP paths; aggregates near the target size
    a = O.gen_g7(24)
    rp, ci = _pattern(a)
    nn = _near_null("const", a.nrows, 1)
    part = PartitionerConfig(8.0, 1.0, 20).build_from_strength(StrengthGraph.new_ls_strength_graph((rp, ci), nn, [1.0], 3))
    part.validate()
    sizes = np.diff(part.agg_ptr)
    assert 6.0 <= part.nnodes() / part.naggs() <= 17.0 and sizes.min() >= 1


@pytest.mark.parametrize("bs", [2, 3])
def test_block_reduce_and_vector_partition_bit_exact(bs):
    """block_size > 1 (partitioners/mod.rs:293-300): LS strength over dofs, aggregated to nodes, self
    loops dropped; the partition is over nodes."""
    import scipy.sparse as sp
    g = O.gen_g7(5, 4, 3).to_scipy()
    a = sp.kron(g, np.ones((bs, bs))).tocsr()
    a.sort_indices()
    rp, ci = a.indptr.astype(np.int64), a.indices.astype(np.int64)
    nn = _near_null("rand", a.shape[0], 3, seed=5)
    w = np.array([1.0, 0.5, 2.0])
    want = OP.block_reduce(OP.new_ls_strength_graph(rp, ci, nn, w, 3), bs)
    sg = StrengthGraph.new_ls_strength_graph((rp, ci), nn, w, 3)
    sg.block_reduce(bs)
    got = _graph_lists(sg)
    assert got == want and len(got) == g.shape[0]
    assert all(j != i for i, nb in enumerate(got) for j, _ in nb)          # filter_diag
    assert max(x for nb in got for _, x in nb) <= 1.0                       # normalised by the global maximum
    n2a, _ = OP.build_partition(rp, ci, nn, w, 4.0, 1.0, 20, block_size=bs)
    part = PartitionerConfig(4.0, 1.0, 20).build_from_strength(sg)
    assert np.array_equal(part.node_to_agg(), n2a) and part.nnodes() == g.shape[0]
