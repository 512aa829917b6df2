import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bigkrls_b200 import _lib
lib = _lib.load(); ctx = _lib.default_context(0)
r = C.c_double()
for S in (1, 2, 4, 8, 16, 32):
    _lib.check(lib.bk_microbench(ctx.handle, 5, S, 5, C.byref(r)))
    print(f"TMA ring, 8 cols x 512 rows per stage, S={S} consecutive row chunks: {r.value:.0f} GB/s", flush=True)
