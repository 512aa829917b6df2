"""Worker for the multi-rank tests: every rank runs the partitioned fit through TorchComm and compares
its column blocks with the oracle (backend nccl on GPUs)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

import krls_oracle as o
from bigkrls_b200 import bigKRLS, crossvalidate_bigKRLS
from bigkrls_b200.dist import TorchComm
from util import relerr

backend = sys.argv[1]
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group(backend, device_id=torch.device("cuda", local))
comm = TorchComm(device=f"cuda:{local}")
X, y = o.synthetic(1500, 6, 77, binary_last=True)
ref = o.bigkrls(y, X, eigtrunc=0.001)
fit = bigKRLS(y, X, eigtrunc=0.001, comm=comm)
c0, c1 = fit["_col_range"]
assert (c0, c1) == (1500 * comm.rank // comm.world, 1500 * (comm.rank + 1) // comm.world)
assert fit["lastkeeper"] == ref["lastkeeper"]
assert abs(fit["lambda"] / ref["lambda"] - 1) < 1e-9
for k in ("coeffs", "yfitted", "derivatives", "avgderivatives", "var.avgderivatives"):
    assert relerr(fit[k], ref[k]) < 1e-8, k
for k in ("K", "vcov.est.c", "vcov.est.fitted"):
    assert fit[k].shape == (1500, c1 - c0)
    assert relerr(fit[k], ref[k][:, c0:c1]) < 1e-8, k
fit.release_device()
# folds sharded over ranks (no data-path collective), statistics gathered
folds = np.random.default_rng(3).permutation(600) % 3 + 1
Xc, yc = o.synthetic(600, 4, 1005)
cv = crossvalidate_bigKRLS(yc, Xc, folds=folds, comm=comm)
rcv = o.crossvalidate_folds(yc, Xc, folds)
for k, v in rcv.items():
    assert relerr(cv[k], v) < 1e-7, k
dist.barrier()
if comm.rank == 0:
    print("DIST_OK")
dist.destroy_process_group()
