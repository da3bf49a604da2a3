import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faer_amg_b200 as F
ctx = F.Context.default(0)
a = F.gallery.diffusion27(ctx, 160)
for which in (0, 2):
    print(which, a.time_kernel(which, 2, 1))
