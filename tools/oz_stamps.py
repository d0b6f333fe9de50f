"""clock64 phase stamps of single CTAs of the tensor-core GEMM kernel (where does the per-tile overhead go?)
   python tools/oz_stamps.py [m n k]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpc_b200._lib import check, lib

m, n, k = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (8192, 8192, 1024)
names = ["start", "setup done", "first operands", "last MMA issued", "C requested", "accumulators done", "stores issued", "end"]
ntiles = (m // 128) * (n // 64)
for cta in (0, 147, 148, 1000, ntiles // 2, ntiles - 200):
    st = (C.c_longlong * 8)()
    check(lib().gpc_bench_oz_stamps(0, m, n, k, cta, st))
    print("CTA %6d: " % cta + "  ".join("%s %.2f us" % (nm, v / 1965.0) for nm, v in zip(names[1:], list(st)[1:])))
