"""Worker for the multi-rank GPU tests (torchrun, one process per GPU, backend nccl for the bootstrap):

  1. self-test of the library's peer-memory collectives (CUDA IPC heaps, flag-synchronised kernels)
  2. the partitioned fit through the NATIVE communicator against the oracle: rank-0 eigensolver (small n),
     distributed dense->band stage (forced at n = 1500, natural at n = 4608), distributed Krylov K X (Neig << N)
  3. the same fit through the generic callback communicator
  4. predict() with newdata rows sharded over the ranks, crossvalidate folds sharded over the ranks

Every rank compares its own column blocks with the oracle at BASELINE.json's tolerances."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

import krls_oracle as o
from bigkrls_b200 import bigKRLS, crossvalidate_bigKRLS, predict
from bigkrls_b200.dist import TorchComm
from util import relerr

backend = sys.argv[1]
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group(backend, device_id=torch.device("cuda", local))
comm = TorchComm(device=f"cuda:{local}")          # native peer communicator
assert comm.peer is not None
bad = comm.selftest()
assert bad == 0, f"peer collectives self-test: {bad} mismatches"
if comm.rank == 0:
    print("PEER_SELFTEST_OK", flush=True)


def check_fit(fit, ref, n, label):
    c0, c1 = fit["_col_range"]
    assert (c0, c1) == (n * comm.rank // comm.world, n * (comm.rank + 1) // comm.world)
    ev, rev = fit["K.eigenvalues"], ref["K.eigenvalues"]
    big = rev >= 1e-3 * rev[0]
    assert np.max(np.abs(ev[big] / rev[big] - 1)) < 1e-9, label
    assert np.max(np.abs(ev - rev)) < 1e-12 * rev[0], label
    assert fit["lastkeeper"] == ref["lastkeeper"], (label, fit["lastkeeper"], ref["lastkeeper"])
    assert abs(fit["lambda"] / ref["lambda"] - 1) < 1e-9, label
    assert fit["_info"]["n_probes"] == ref["_nprobe"], label
    for k in ("coeffs", "yfitted", "derivatives", "avgderivatives", "var.avgderivatives"):
        assert relerr(fit[k], ref[k]) < 1e-8, (label, k, relerr(fit[k], ref[k]))
    for k in ("K", "vcov.est.c", "vcov.est.fitted"):
        assert fit[k].shape == (n, c1 - c0)
        assert relerr(fit[k], ref[k][:, c0:c1]) < 1e-8, (label, k)
    for k in ("Looe", "Neffective", "R2", "R2AME"):
        assert abs(fit[k] / ref[k] - 1) < 1e-8, (label, k)
    if comm.rank == 0:
        i = fit["_info"]
        print("DIST_CASE_OK %s: lambda %.9f lastkeeper %d | t_total %.4f eigen %.4f (sy2sb %.4f sb2st %.4f dc %.4f bt %.4f) "
              "kernel %.4f lambda %.4f coef %.4f vcov %.4f deriv %.4f" %
              (label, fit["lambda"], fit["lastkeeper"], i["t_total"], i["t_eigen"], i["t_sy2sb"], i["t_sb2st"], i["t_dc"],
               i["t_backtransform"], i["t_kernel"], i["t_lambda"], i["t_coef"], i["t_vcov"], i["t_deriv"]), flush=True)


# ---- 2a. small n: eigensolver on rank 0, every exchange through the peer communicator ---------------------------
X, y = o.synthetic(1500, 6, 77, binary_last=True)
ref = o.bigkrls(y, X, eigtrunc=0.001)
fit = bigKRLS(y, X, eigtrunc=0.001, comm=comm)
check_fit(fit, ref, 1500, "native/rank0-eigen n=1500")
# sharded predict against the oracle (rows of newdata split over the ranks, gathered)
Xn, _ = o.synthetic(101, 6, 78, binary_last=True)
rp = o.predict(ref, Xn, se_pred=True)
gp = predict(fit, Xn, se_pred=True, comm=comm)
assert relerr(gp["predicted"], rp["predicted"]) < 1e-8
assert relerr(gp["se.pred"], rp["se.pred"]) < 1e-7
assert relerr(gp["vcov.est.pred"], rp["vcov.est.pred"]) < 1e-7
assert relerr(gp["newdataK"], rp["newdataK"]) < 1e-13
fit.release_device()
# ---- 2b. the distributed dense->band stage, forced at the same size (odd n: ragged last block) ---------------
os.environ["BK_EIG_TWOSTAGE"] = "1"
for n_, seed in ((1500, 77), (1531, 79)):
    Xb, yb = o.synthetic(n_, 6, seed, binary_last=True)
    refb = ref if n_ == 1500 else o.bigkrls(yb, Xb, eigtrunc=0.001)
    fit = bigKRLS(yb, Xb, eigtrunc=0.001, comm=comm)
    check_fit(fit, refb, n_, "native/distributed sy2sb n=%d (forced)" % n_)
    fit.release_device()
del os.environ["BK_EIG_TWOSTAGE"]
# ---- 2c. natural two-stage size -------------------------------------------------------------------------
X2, y2 = o.synthetic(4608, 8, 81)
ref2 = o.bigkrls(y2, X2)
fit = bigKRLS(y2, X2, comm=comm)
check_fit(fit, ref2, 4608, "native/distributed sy2sb n=4608")
fit.release_device()
# ---- 2d. Neig << N: block Krylov with K partitioned over the ranks ------------------------------------------
X3, y3 = o.synthetic(2400, 8, 7)
kw = dict(Neig=300, eigtrunc=0.001, which_derivatives=[1, 3, 5])
ref3 = o.bigkrls(y3, X3, **kw)
fit = bigKRLS(y3, X3, comm=comm, **kw)
check_fit(fit, ref3, 2400, "native/distributed Krylov n=2400 Neig=300")
assert fit["_info"]["krylov_matvecs"] > 0
fit.release_device()

# ---- 3. the generic callback communicator (NCCL through torch.distributed) -----------------------------------
comm_cb = TorchComm(device=f"cuda:{local}", native=False)
fit = bigKRLS(y, X, eigtrunc=0.001, comm=comm_cb)
check_fit(fit, ref, 1500, "callbacks n=1500")
fit.release_device()

# ---- 4. folds sharded over ranks (no data-path collective), statistics gathered ---------------------------------
folds = np.random.default_rng(3).permutation(600) % 3 + 1
Xc, yc = o.synthetic(600, 4, 1005)
cv = crossvalidate_bigKRLS(yc, Xc, folds=folds, comm=comm)
rcv = o.crossvalidate_folds(yc, Xc, folds)
for k, v in rcv.items():
    assert relerr(cv[k], v) < 1e-7, k
# Ncores: a 1-GPU fit inside the multi-rank job (ranks >= Ncores return None)
f1 = bigKRLS(y, X, eigtrunc=0.001, comm=comm, Ncores=1)
if comm.rank == 0:
    assert f1["_col_range"] == (0, 1500) and relerr(f1["coeffs"], ref["coeffs"]) < 1e-8
else:
    assert f1 is None
dist.barrier()
comm.close()
if comm.rank == 0:
    print("DIST_OK", flush=True)
dist.destroy_process_group()
