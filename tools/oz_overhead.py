import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpc_b200 as G
from gpc_b200._lib import check, lib
L = lib(); ms = C.c_double(0)
for (m, n, k) in [(8192, 8192, 128), (8192, 8192, 256), (8192, 8192, 512), (8192, 8192, 1024)]:
    check(L.gpc_bench_gemm(0, m, n, k, 0, 0, 0, 108, 5, C.byref(ms)))
    tiles = (m // 128) * (n // 64)
    print("m=%d n=%d k=%d: %.1f us total, %.2f us per tile-wave (%d waves)" % (m, n, k, ms.value * 1e3, ms.value * 1e3 / (tiles / 148.0), tiles // 148))
