"""tools/prof_syrk.py -- run the SYRK trailing update (DMMA GEMM engine, lower-only) in isolation, for ncu."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpc_b200 as G  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
k = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
ms = C.c_double(0)
G._lib.check(G.lib().gpc_bench_syrk(0, n, k, 2, C.byref(ms)))
print("syrk n=%d k=%d: %.3f ms, %.2f TFLOP/s (lower tiles)" % (n, k, ms.value, n * (n + 128) * k / ms.value / 1e9))
