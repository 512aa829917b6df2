"""Time one fused fit at (N, P) on the GPU and print the per-stage break-down."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np  # noqa: E402

from bigkrls_b200 import bigKRLS  # noqa: E402
import krls_oracle as o  # noqa: E402

N, P = int(sys.argv[1]), int(sys.argv[2])
kw = {}
if len(sys.argv) > 3:
    kw["eigtrunc"] = float(sys.argv[3])
reps = 1 if (len(sys.argv) > 4 and sys.argv[4] == "once") else int(os.environ.get("FIT_REPS", "2"))
if len(sys.argv) > 5:
    kw["Neig"] = int(sys.argv[5])
if len(sys.argv) > 6:
    kw["which_derivatives"] = [int(v) for v in sys.argv[6].split(",")]
X, y = o.synthetic(N, P, 1003 if (N, P) == (20000, 10) else 1000 + P)
for rep in range(reps):
    t0 = time.time()
    fit = bigKRLS(y, X, return_squares=False, **kw)
    t1 = time.time()
    info = fit["_info"]
    fit.release_device()
    print(json.dumps({"N": N, "P": P, "wall": t1 - t0, **{k: (round(v, 6) if isinstance(v, float) else v)
                                                       for k, v in info.items()}}))
