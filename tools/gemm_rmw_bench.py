"""Short-k read-modify-write DGEMM shapes of the tridiagonalisation updates (C -= [V W][W V]')."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bigkrls_b200 import _lib  # noqa: E402

lib = _lib.load()
ctx = _lib.default_context(0)
r = C.c_double()
for (m, k, lower, beta) in [(16384, 128, 0, 1.0), (16384, 128, 1, 1.0), (16384, 128, 2, 1.0), (16384, 128, 0, 0.0),
                            (8192, 128, 2, 1.0), (16384, 256, 2, 1.0), (16384, 2048, 1, 1.0)]:
    best = 1e9
    for _ in range(3):
        _lib.check(lib.bk_dgemm_bench(ctx.handle, 0, 1, m, m, k, lower, beta, 5, C.byref(r)))
        best = min(best, r.value)
    fl = 2.0 * m * m * k * (0.5 if lower else 1.0)
    print(f"NT m=n={m} k={k} lower={lower} beta={beta}: {best*1e3:.3f} ms  {fl/best*1e-12:.1f} TF/s (useful)")
