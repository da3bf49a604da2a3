"""Few launches of each fused fine-level kernel (for ncu --set full): SpMV, residual, smoother sweep."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_amg_b200 as F
ctx = F.Context.default(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
a = F.gallery.poisson7(ctx, n)
for which in (0, 1, 2):
    print(which, a.time_kernel(which, 2, 1))
