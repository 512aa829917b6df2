"""torchrun worker: headline fit on N ranks with the per-segment profile of the distributed dense->band stage
(BK_DIST_PROF=1 prints it on every rank's stderr).  tools/r2_run10.sh"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from bigkrls_b200 import bigKRLS, _lib
from bigkrls_b200.dist import TorchComm
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = TorchComm(device=f"cuda:{local}", ctx=_lib.default_context(local))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rng = np.random.default_rng(1003)
X = rng.standard_normal((N, 10)); y = np.sin(X[:, 0]) + X[:, 1] * X[:, 2] + 0.5 * rng.standard_normal(N)
for rep in range(3):
    os.environ.pop("BK_DIST_PROF", None)
    if rep == 2 and comm.rank in (0, 1):
        os.environ["BK_DIST_PROF"] = "1"
    fit = bigKRLS(y, X, eigtrunc=0.001, comm=comm, return_squares=False)
    i = fit["_info"]
    fit.release_device()
    if comm.rank == 0:
        print(json.dumps({k: round(i[k], 5) for k in ("t_total", "t_eigen", "t_sy2sb", "t_sb2st", "t_dc", "t_backtransform")}), flush=True)
comm.close()
dist.destroy_process_group()
