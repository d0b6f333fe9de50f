"""tools/gemm_sweep.py -- time the DMMA GEMM engine per tile configuration on the shapes the recursion produces."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpc_b200 as G  # noqa: E402

L = G.lib()
shapes = [  # (m, n, k, a_kc, b_kc, lower, label)
    (16384, 16384, 4096, 0, 0, 1, "syrk 16384 k4096"),
    (8192, 8192, 2048, 0, 0, 1, "syrk 8192 k2048"),
    (4096, 4096, 4096, 0, 0, 1, "syrk 4096 k4096"),
    (2048, 2048, 2048, 0, 0, 1, "syrk 2048 k2048"),
    (1024, 1024, 1024, 0, 0, 1, "syrk 1024 k1024"),
    (512, 512, 512, 0, 0, 1, "syrk 512 k512"),
    (4096, 4096, 4096, 0, 1, 0, "gemm NN 4096^3"),
    (4096, 4096, 4096, 1, 1, 0, "gemm TN 4096^3"),
    (2048, 2048, 2048, 0, 1, 0, "gemm NN 2048^3"),
    (4096, 2048, 2048, 0, 0, 0, "trsm-upd 4096x2048 k2048"),
    (4096, 512, 512, 0, 0, 0, "trsm-upd 4096x512 k512"),
    (4096, 128, 128, 0, 0, 0, "trsm-upd 4096x128 k128"),
    (1024, 512, 512, 0, 1, 0, "NN 1024x512 k512"),
    (256, 256, 256, 0, 0, 0, "256^3"),
]
print("%-28s %s" % ("shape", "  ".join("cfg%d: ms (TF/s)" % c for c in range(5))))
for m, n, k, akc, bkc, lower, label in shapes:
    fl = (m * (m + 128) * k) if lower else 2.0 * m * n * k
    out = []
    for cfg in range(5):
        ms = C.c_double(0)
        reps = 3 if fl > 1e10 else 20
        rc = L.gpc_bench_gemm(0, m, n, k, akc, bkc, lower, cfg, reps, C.byref(ms))
        out.append("%8.3f (%5.2f)" % (ms.value, fl / ms.value / 1e9) if rc == 0 else "   error: %s" % L.gpc_last_error().decode()[:30])
    print("%-28s %s" % (label, "  ".join(out)), flush=True)
