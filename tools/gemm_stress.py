import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bigkrls_b200 import _lib
from bigkrls_b200._lib import check, dptr
lib = _lib.load(); ctx = _lib.default_context(0)
rng = np.random.default_rng(1)
for (m, k, ld) in [(3036, 128, 3100), (3936, 128, 4000), (2736, 128, 2800), (3008, 128, 3072)]:
    A = np.asfortranarray(rng.standard_normal((ld, k)))
    B = np.asfortranarray(rng.standard_normal((ld, k)))
    C0 = np.asfortranarray(rng.standard_normal((ld, m)))
    off = ld - m
    ref = C0[off:, :] - A[off:] @ B[off:].T
    outs = []
    for rep in range(6):
        Cc = C0.copy(order="F")
        Cv = Cc[off:, :]
        check(lib.bk_debug_gemm(ctx.handle, 0, 1, m, m, k, -1.0,
                                A[off:].ctypes.data_as(_lib.c_double_p), ld, B[off:].ctypes.data_as(_lib.c_double_p), ld,
                                1.0, Cv.ctypes.data_as(_lib.c_double_p), ld, 1, 3))
        outs.append(Cv.copy())
    low = np.tril(np.ones((m, m), bool))
    errs = [np.max(np.abs((o - ref)[low])) for o in outs]
    r2r = max(np.max(np.abs((outs[0] - o)[low])) for o in outs[1:])
    print(f"gemm lower m={m} k={k} ld={ld}: err {max(errs):.2e} run2run {r2r:.2e}", flush=True)
