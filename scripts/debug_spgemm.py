import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle as O
import faer_amg_b200 as F
from util import random_csr, to_dev
ctx = F.Context.default(0)
rng = np.random.default_rng(7)
random_csr(rng, 300, 200, rng.integers(0, 6, 300)); random_csr(rng, 200, 250, rng.integers(0, 8, 200))
oa = random_csr(rng, 120, 300, rng.integers(10, 30, 120)); ob = random_csr(rng, 300, 400, rng.integers(5, 30, 300))
c = to_dev(ctx, oa) @ to_dev(ctx, ob)
oc = O.spgemm(oa, ob)
rp, ci, v = c.to_host(); rp = rp.astype(np.int64); ci = ci.astype(np.int64)
print("nnz", c.nnz, oc.nnz, "rp equal", np.array_equal(rp, oc.row_ptr))
lens_d, lens_o = np.diff(rp), np.diff(oc.row_ptr)
bad = np.nonzero(lens_d != lens_o)[0]
print("rows with different length:", bad[:20], lens_d[bad[:20]], lens_o[bad[:20]])
ub = np.array([sum(ob.row_ptr[k+1]-ob.row_ptr[k] for k in oa.col[oa.row_ptr[i]:oa.row_ptr[i+1]]) for i in range(oa.nrows)])
print("ub range", ub.min(), ub.max(), "classes", np.bincount(np.digitize(ub, [129, 1025, 8193])))
if len(bad) == 0:
    for i in range(oa.nrows):
        a, b = ci[rp[i]:rp[i+1]], oc.col[oc.row_ptr[i]:oc.row_ptr[i+1]]
        if not np.array_equal(a, b):
            print("row", i, "ub", ub[i], "dev", a[:40], "orc", b[:40], "sorted?", np.all(np.diff(a) > 0), "same set", set(a) == set(b))
            break
else:
    print("ub of bad rows", ub[bad[:20]])
