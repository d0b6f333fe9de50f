"""tools/prof_gemm.py m n k lower cfg [reps] -- one GEMM shape on device scratch (cfg: -1 heuristic DMMA, 0..4 DMMA tile
config, 100+S Ozaki/tcgen05 with S slices), for ncu captures and quick timings."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpc_b200 as G  # noqa: E402
from gpc_b200._lib import check, lib  # noqa: E402

m, n, k, lower, cfg = [int(x) for x in sys.argv[1:6]]
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 1
ms = C.c_double(0)
check(lib().gpc_bench_gemm(0, m, n, k, 0, 0, lower, cfg, reps, C.byref(ms)))
fl = (m * (m + 128) * k) if lower else 2.0 * m * n * k
print("gemm m=%d n=%d k=%d lower=%d cfg=%d: %.3f ms  %.1f TFLOP/s-equivalent" % (m, n, k, lower, cfg, ms.value, fl / ms.value / 1e9))
