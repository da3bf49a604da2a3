"""BASELINE.json configurations at their full sizes on the GPU against the committed oracle known answers
(tests/golden/oracle_golden.json: level rows / non-zeros of the hierarchy, PCG iteration count, norm and sum of the
solution) -- the oracle itself is not run here (minutes of CPU time; tests/golden/make_golden*.py regenerates them).
    configs[1]  G7 128^3        g7_128_l1        configs[2]  G27 192^3 (and 96^3)   g27_192_l1, g27_96_l1
    configs[3]  G7 256^3        g7_256_l1        (multi-GPU runs of [3]/[4]: scripts/dist_check.py, bench.py --gpus N)"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.json")))["amg"]


@pytest.mark.parametrize("case", ["g7_128_l1", "g27_96_l1", "g27_192_l1", "g7_256_l1"])
def test_full_size_config_against_oracle_golden(ctx, case):
    import faer_amg_b200 as F
    g = GOLD[case]
    dims = tuple(g["dims"])
    a = (F.gallery.diffusion27 if case.startswith("g27") else F.gallery.poisson7)(ctx, *dims)
    n = a.nrows
    nn = np.full((n, 1), 1.0 / np.sqrt(n))
    h = F.HierarchyConfig(1000, F.AggregationConfig(1, 1, F.GeometricPartitioner(dims))).build(F.SparseMatOp(a), nn)
    assert h.levels() == g["levels"]
    assert [m.mat_ref().nrows for m in h.operators()] == g["level_rows"]
    assert [m.mat_ref().nnz for m in h.operators()] == g["level_nnz"]          # bit-exact sparsity pattern sizes, every level
    assert abs(h.op_complexity() - g["op_complexity"]) < 1e-12
    mg = F.MultigridConfig(smoother="l1").build(h)
    want = g["iters"]["1e-08"]
    x = np.zeros(n)
    info = F.conjugate_gradient(x, mg, a, np.ones(n), F.CgParams(0.0, 1e-8, 1000))
    assert abs(info.iter_count - want["iters"]) <= 1                            # north_star: +-1
    # the solution itself: norm and sum agree with the oracle's far below the solve tolerance
    assert abs(np.linalg.norm(x) - want["x_norm"]) <= 1e-7 * want["x_norm"]
    assert abs(x.sum() - want["x_sum"]) <= 1e-7 * abs(want["x_sum"])
    if info.iter_count == want["iters"]:
        assert abs(info.rel_residual - want["rel_residual"]) <= 1e-3 * want["rel_residual"]
