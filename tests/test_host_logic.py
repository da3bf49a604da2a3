"""CPU tests of host-side logic: partitions, row splits for the multi-GPU path, and the
world_size-2 rendezvous plumbing over gloo."""
import os
import socket
import sys

import numpy as np
import pytest

import oracle as O
from faer_amg_b200.partitioners import GeometricPartitioner, Partition, geometric_partition
from faer_amg_b200.distributed import level_row_splits, slab_splits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_geometric_partition_matches_oracle_aggregates():
    for dims in [(8, 8, 8), (6, 4, 2), (5, 7, 3), (16, 1, 1)]:
        part, coarse = geometric_partition(dims)
        ap, an, dc = O.geometric_aggregates(dims)
        assert np.array_equal(part.agg_ptr, ap) and np.array_equal(part.agg_nodes, an) and tuple(coarse) == tuple(dc)
        for agg in part.aggregates():
            assert np.all(np.diff(agg) > 0)
        assert np.array_equal(np.sort(part.agg_nodes), np.arange(np.prod(dims)))


def test_partition_validation():
    with pytest.raises(ValueError):
        Partition([0, 2, 3], [0, 1, 1], 3)
    p = Partition.from_node_to_agg([1, 0, 1, 0])
    assert p.naggs() == 2 and p.node_to_agg().tolist() == [1, 0, 1, 0]


def test_geometric_partitioner_tracks_levels():
    gp = GeometricPartitioner((8, 8, 8))
    assert gp(0, None, None).naggs() == 64
    assert gp(1, None, None).naggs() == 8
    assert gp.dims == [(8, 8, 8), (4, 4, 4), (2, 2, 2)]


def test_slab_splits_keep_aggregates_on_one_rank():
    # z-slabs of even thickness: 2x2x2 aggregates never straddle a rank boundary
    dims = [(16, 16, 16), (8, 8, 8), (4, 4, 4), (2, 2, 2)]
    for nranks in (1, 2, 4, 8):
        splits = level_row_splits(dims, nranks)
        assert len(splits) == len(dims)
        for lvl, (d, s) in enumerate(zip(dims, splits)):
            assert s[0] == 0 and s[-1] == np.prod(d) and np.all(np.diff(s) >= 0)
        fine, coarse = splits[0], splits[1]
        part, _ = geometric_partition(dims[0])
        n2a = part.node_to_agg()
        for r in range(nranks):
            aggs = np.unique(n2a[fine[r]:fine[r + 1]])
            if len(aggs):
                assert aggs.min() >= coarse[r] and aggs.max() < coarse[r + 1]
    assert slab_splits((4, 4, 6), 4).tolist() == [0, 32, 64, 96, 96] or slab_splits((4, 4, 6), 4)[-1] == 96


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from faer_amg_b200.distributed import broadcast_unique_id, gather_rows, level_row_splits
    # rank 0 invents the 128-byte id (no NCCL needed for the plumbing test); everyone must agree
    uid = broadcast_unique_id(make=lambda: bytes(range(128)))
    splits = level_row_splits([(4, 4, 8), (2, 2, 4)], world)
    lo, hi = splits[0][rank], splits[0][rank + 1]
    full = gather_rows(np.arange(lo, hi, dtype=np.float64), splits[0])
    q.put((rank, uid == bytes(range(128)), bool(np.array_equal(full, np.arange(128.0)))))
    dist.destroy_process_group()


def test_world_size_2_plumbing_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True, True), (1, True, True)]


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` is the CPU arm of the contract: it must run without a GPU and
    print one JSON line with the reference keys."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--grid", "16", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "ms" and line["higher_is_better"] is False
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0
    assert line["config"]["grid"] == [16, 16, 16]


def test_utils_convergence_factor_and_symmetry_test_logic():
    """utils.rs:691-736 / multigrid.rs:520-580 restated over the LinOp::apply surface: checked here with
    numpy operators (no device): ||E||_A of a Jacobi iteration equals the spectral radius of I - D^-1 A."""
    from faer_amg_b200.utils import approx_convergence_factor, symmetry_test

    a = O.gen_g1(33).to_scipy().toarray()
    n = a.shape[0]
    d = 0.66 / np.diag(a)

    class Op:
        nrows = n

        def __init__(self, m):
            self.m = m

        def apply(self, x):
            return self.m @ x

    cf = approx_convergence_factor(Op(a), Op(np.diag(d)), iterations=400, test_vecs=3, seed=1)
    rho = np.max(np.abs(np.linalg.eigvals(np.eye(n) - np.diag(d) @ a)))
    assert abs(cf - rho) < 1e-3 * rho
    err, rel = symmetry_test(Op(np.linalg.inv(a)), test_dim=4, seed=2)
    assert err < 1e-9
    ns = a.copy(); ns[0, 5] += 1.0
    assert symmetry_test(Op(ns), test_dim=4, seed=2)[0] > 1e-3


def test_dist_geometric_partitioner_matches_the_global_partition():
    """Every rank's local boxes, concatenated in rank order, are the global boxes (aggregates never
    straddle ranks), level after level."""
    from faer_amg_b200.distributed import DistGeometricPartitioner, fine_plane_splits
    from faer_amg_b200.partitioners import geometric_partition
    for dims, nranks in [((8, 6, 16), 4), ((5, 7, 24), 3), ((4, 4, 64), 8), ((6, 6, 9), 1)]:
        dp = DistGeometricPartitioner(dims)
        rs = fine_plane_splits(dims, nranks)
        level = 0
        while True:
            d = dp.level_dims(level)
            gpart, coarse = geometric_partition(d)
            plane = d[0] * d[1]
            ptr, nodes, nxt = [0], [], [0]
            try:
                for r in range(nranks):
                    if rs[r + 1] == rs[r]:
                        nxt.append(nxt[-1]); continue
                    lp = dp.local(level, int(rs[r]), int(rs[r + 1]))
                    nodes.append(lp.agg_nodes + int(rs[r]))
                    ptr.extend((lp.agg_ptr[1:] + len(np.concatenate(nodes[:-1])) if len(nodes) > 1 else lp.agg_ptr[1:]).tolist())
                    nxt.append(nxt[-1] + lp.naggs())
            except ValueError:
                break  # slabs no longer hold whole boxes: this level would be replicated
            assert np.array_equal(np.asarray(ptr), gpart.agg_ptr), (dims, nranks, level)
            assert np.array_equal(np.concatenate(nodes), gpart.agg_nodes), (dims, nranks, level)
            rs = np.asarray(nxt, dtype=np.int64)
            level += 1
            if int(np.prod(coarse)) <= 8:
                break
        assert level >= 1, (dims, nranks)
