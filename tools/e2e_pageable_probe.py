"""e2e of bigKRLS(y, X) at the headline size into (a) library-pinned buffers, (b) plain pageable numpy memory,
(c) np.memmap files in /dev/shm - the stand-in for the reference's shared-memory / file-backed big.matrix outputs."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bigkrls_b200 import bigKRLS, _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rng = np.random.default_rng(1003)
X = rng.standard_normal((N, 10)); y = np.sin(X[:, 0]) + X[:, 1] * X[:, 2] + 0.5 * rng.standard_normal(N)
ctx = _lib.default_context(0)
def shm_alloc(name, shape):
    path = "/dev/shm/bk_%s_%d.bin" % (name.replace(".", "_"), os.getpid())
    return np.memmap(path, dtype=np.float64, mode="w+", shape=shape, order="F")
_pre = {}
def shm_prefaulted(name, shape):
    """The reference's R wrappers allocate every output big.matrix with init = 0 before the native call
    (R/bigKRLS_Rcpp_functions.R:206,235,251): the pages exist when the library writes into them."""
    if name not in _pre:
        _pre[name] = shm_alloc("pre_" + name, shape)
        _pre[name][:] = 0.0
    return _pre[name]
res = {}
for label, kw in (("pinned", dict(pinned=True)), ("pageable", dict()), ("shm_memmap", dict(squares_alloc=shm_alloc)),
                  ("shm_memmap_prefaulted", dict(squares_alloc=shm_prefaulted))):
    ts = []
    for rep in range(3):
        t0 = time.perf_counter()
        fit = bigKRLS(y, X, eigtrunc=0.001, ctx=ctx, **kw)
        ts.append(time.perf_counter() - t0)
        chk = float(fit["K"][123, 456] + fit["vcov.est.c"][7, 9] + fit["vcov.est.fitted"][N - 1, N - 2])
        fit.release_device()
        if kw.get("pinned"): fit.release_pinned()
        del fit
    res[label] = {"seconds": ts, "check": chk}
for f in os.listdir("/dev/shm"):
    if f.startswith("bk_"): os.unlink(os.path.join("/dev/shm", f))
print(json.dumps(res))
