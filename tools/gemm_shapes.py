"""tools/gemm_shapes.py -- DMMA vs Ozaki(8) time for a list of GEMM shapes (engine cost-model calibration)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpc_b200 as G  # noqa: E402
from gpc_b200._lib import check, lib  # noqa: E402

L = lib()
ms = C.c_double(0)
shapes = [(1024, 1024, 1024, 0), (1024, 1024, 2048, 0), (2048, 2048, 512, 0), (2048, 2048, 1024, 0), (2048, 2048, 2048, 0),
          (2048, 2048, 2048, 1), (4096, 4096, 1024, 1), (4096, 1024, 1024, 0), (8192, 512, 512, 0), (8192, 1024, 1024, 0),
          (16384, 1024, 1024, 0), (16384, 2048, 2048, 0), (4096, 4096, 4096, 1), (4096, 4096, 2048, 0), (2048, 4096, 4096, 0)]
for (m, n, k, lower) in shapes:
    out = []
    for cfg in (-1, 108):
        check(L.gpc_bench_gemm(0, m, n, k, 0, 0, lower, cfg, 10, C.byref(ms)))
        out.append(ms.value * 1e3)
    print("m=%5d n=%5d k=%5d lower=%d: DMMA %8.1f us  Ozaki8 %8.1f us  ratio %.2f" % (m, n, k, lower, out[0], out[1], out[0] / out[1]), flush=True)
