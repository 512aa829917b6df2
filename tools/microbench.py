"""FP64 roofline denominators on this box: DFMA pipe, DMMA (m8n8k4) pipe, HBM copy, and the
library DGEMM at a few shapes; cuBLAS DGEMM as a comparator only (never on the product path)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bigkrls_b200 import _lib  # noqa: E402

lib = _lib.load()
ctx = _lib.default_context(0)
out = {"device": ctx.device_info()}
r = C.c_double()
for kind, name in ((0, "dfma_tflops"), (1, "dmma_tflops"), (2, "hbm_copy_gbs"), (3, "hbm_read_ldg_gbs"),
                   (4, "hbm_read_tma_ring_gbs")):
    _lib.check(lib.bk_microbench(ctx.handle, kind, 0, 0, C.byref(r)))
    out[name] = r.value
for which, name in ((0, "dfma"), (1, "dmma884"), (2, "dadd"), (3, "shfl64_dadd"), (4, "smem_roundtrip_dadd")):
    _lib.check(lib.bk_microbench(ctx.handle, 7, which, 0, C.byref(r)))
    out.setdefault("dependent_chain_latency_cycles", {})[name] = r.value
if "--peaks-only" in sys.argv:
    print(json.dumps(out, indent=1))
    sys.exit(0)
shapes = [(0, 1, 8192, 8192, 8192, 0), (0, 0, 8192, 8192, 8192, 0), (1, 0, 8192, 8192, 8192, 0),
          (0, 1, 16384, 16384, 2048, 1), (0, 1, 16384, 16384, 128, 1), (0, 0, 20000, 22, 20000, 0),
          (1, 0, 2000, 10, 20000, 0), (0, 0, 16384, 2048, 64, 0), (1, 0, 64, 2048, 16384, 0)]
out["dgemm"] = []
for ta, tb, m, n, k, lower in shapes:
    _lib.check(lib.bk_dgemm_bench(ctx.handle, ta, tb, m, n, k, lower, 0.0, 3, C.byref(r)))
    flops = 2.0 * m * n * k * (0.5 if lower else 1.0)
    out["dgemm"].append({"ta": ta, "tb": tb, "m": m, "n": n, "k": k, "lower": lower, "sec": r.value,
                         "tflops": flops / r.value * 1e-12})
try:
    import torch
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    out["cublas_dgemm_8192_tflops_comparator"] = 2 * 8192 ** 3 / best * 1e-12
except Exception as e:  # noqa: BLE001
    out["cublas_error"] = str(e)
print(json.dumps(out, indent=1))
