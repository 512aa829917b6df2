import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
from bigkrls_b200 import bigKRLS, _lib
import krls_oracle as o
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
X, y = o.synthetic(N, 10, 1003)
ctx = _lib.default_context(0)
for rep in range(3):
    t0 = time.perf_counter()
    if rep == 2:
        pr = cProfile.Profile(); pr.enable()
    fit = bigKRLS(y, X, eigtrunc=0.001, pinned=True, ctx=ctx)
    if rep == 2:
        pr.disable()
    t1 = time.perf_counter()
    print("rep", rep, "wall", round(t1 - t0, 3), "t_total", round(fit["_info"]["t_total"], 3))
    fit.release_device(); fit.release_pinned()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
