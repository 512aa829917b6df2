"""DGEMM shapes of the two-stage reduction under forced tile configurations (tuning aid).
usage: gemm_shape_bench.py  (reads BK_GEMM_FORCE from the environment)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bigkrls_b200 import _lib  # noqa: E402

lib = _lib.load()
ctx = _lib.default_context(0)
r = C.c_double()
shapes = [  # ta tb m n k lower beta
    (0, 0, 16384, 64, 16384, 0, 0.0),    # Z = A22 (V T)
    (0, 0, 8192, 64, 8192, 0, 0.0),
    (0, 1, 16384, 16384, 128, 2, 1.0),   # rank-128 update, mirrored
    (0, 1, 16384, 16384, 128, 0, 0.0),
    (0, 1, 8192, 8192, 8192, 0, 0.0),
    (0, 0, 8192, 8192, 8192, 0, 0.0),
    (1, 0, 64, 286, 16384, 0, 0.0),      # V' Z (q1)
    (0, 0, 16384, 286, 64, 0, 1.0),      # Z -= V W (q1)
]
for (ta, tb, m, n, k, lower, beta) in shapes:
    best = 1e9
    for _ in range(3):
        _lib.check(lib.bk_dgemm_bench(ctx.handle, ta, tb, m, n, k, lower, beta, 5, C.byref(r)))
        best = min(best, r.value)
    fl = 2.0 * m * n * k * (0.5 if lower else 1.0)
    print(f"{'T' if ta else 'N'}{'T' if tb else 'N'} m={m} n={n} k={k} lower={lower} beta={beta}: {best*1e3:.3f} ms  {fl/best*1e-12:.1f} TF/s")
