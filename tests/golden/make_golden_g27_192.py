"""Adds the BASELINE configs[2] known answer (27-point anisotropic diffusion, 192^3, geometric 2x2x2 aggregates,
L1 smoother, rel 1e-8) to oracle_golden.json from the CPU oracle.  ~20-30 minutes and ~20 GB on 8 cores."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import oracle as O  # noqa: E402
from make_golden import amg_case  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
path = os.path.join(HERE, "oracle_golden.json")
case = amg_case(O.gen_g27, (192, 192, 192), "l1", tols=(1e-8,))
gold = json.load(open(path))
gold["amg"]["g27_192_l1"] = case
json.dump(gold, open(path, "w"), indent=1, sort_keys=True)
print("g27_192_l1", case["iters"], case["level_nnz"], f'{case["seconds"]:.0f}s')
