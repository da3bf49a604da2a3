"""Regenerates the committed known-answer fixtures from the CPU oracle.

The reference ships no tests or golden vectors and cannot be built here, so these
fixtures are *oracle* outputs: they pin the restatement against regressions and give the GPU tests
full-size known answers without the oracle having to run at those sizes in CI.

    python tests/golden/make_golden.py [--big]     # --big adds G7(128) and G7(256) (minutes, GBs)
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
EPS = float(np.finfo(float).eps)


def simple_geometric(max_refinement=10):
    """examples/simple_geometric.rs:176-301: DOFs | PCG+Jacobi | PCG+MG | Stat+MG."""
    rows = []
    for refinement in range(2, max_refinement + 1):
        ne = 10 * 2 ** refinement
        a = O.gen_g1(ne)
        mg = O.Multigrid()
        mg.add_level(a, O.new_jacobi(a, 0.66))
        for level in range(1, refinement + 1):
            ce = 10 * 2 ** (refinement - level)
            m = O.gen_g1(ce)
            sm = "cholesky" if level == refinement else O.new_jacobi(m, 0.66)
            mg.add_level(m, sm, O.gen_g1_restrict(ce - 1), O.gen_g1_interp(ce - 1))
        b = np.ones(ne - 1)
        _, i1 = O.pcg(a, b, O.new_jacobi(a, 0.66), rel_tol=1e-8, abs_tol=EPS, max_iters=6000)
        _, i2 = O.pcg(a, b, mg, rel_tol=1e-8, abs_tol=EPS, max_iters=6000)
        _, i3 = O.stationary_solver(a, b, mg, 6000, 1e-8)
        rows.append({"dofs": ne - 1, "pcg_jacobi": i1.iters, "pcg_mg": i2.iters, "stat_mg": i3})
    return rows


def amg_case(gen, dims, smoother, tols=(1e-8, 1e-12)):
    t0 = time.time()
    a = gen(*dims)
    nn = np.full((a.nrows, 1), 1.0 / np.sqrt(a.nrows))
    h = O.build_hierarchy(a, nn, dims)
    mg = O.multigrid_from_hierarchy(h, smoother)
    out = {"dims": list(dims), "smoother": smoother, "levels": h.levels,
           "level_rows": [o.nrows for o in h.operators], "level_nnz": [o.nnz for o in h.operators],
           "op_complexity": h.op_complexity(), "iters": {}}
    b = np.ones(a.nrows)
    for tol in tols:
        x, info = O.pcg(a, b, mg, rel_tol=tol, abs_tol=0.0, max_iters=1000)
        out["iters"][f"{tol:g}"] = {"iters": info.iters, "rel_residual": info.rel_residual, "status": info.status,
                                    "x_norm": float(np.linalg.norm(x)), "x_sum": float(x.sum())}
    out["seconds"] = time.time() - t0
    return out


def main():
    big = "--big" in sys.argv
    gold = {"simple_geometric": simple_geometric(), "amg": {}}
    cases = [("g7_16", O.gen_g7, (16, 16, 16)), ("g7_32", O.gen_g7, (32, 32, 32)), ("g7_64", O.gen_g7, (64, 64, 64)),
             ("g7_48x32x16", O.gen_g7, (48, 32, 16)), ("g27_24", O.gen_g27, (24, 24, 24))]
    if big:
        cases += [("g7_128", O.gen_g7, (128, 128, 128)), ("g27_96", O.gen_g27, (96, 96, 96))]
    if "--huge" in sys.argv:
        cases += [("g7_256", O.gen_g7, (256, 256, 256))]
    path = os.path.join(HERE, "oracle_golden.json")
    if os.path.exists(path):
        old = json.load(open(path))
        gold["amg"].update(old.get("amg", {}))
    for name, gen, dims in cases:
        for sm in ("l1", "jacobi"):
            key = f"{name}_{sm}"
            gold["amg"][key] = amg_case(gen, dims, sm)
            print(key, gold["amg"][key]["iters"], f'{gold["amg"][key]["seconds"]:.1f}s', flush=True)
            json.dump(gold, open(path, "w"), indent=1, sort_keys=True)
    json.dump(gold, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
