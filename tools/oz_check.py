"""tools/oz_check.py -- bring-up of the Ozaki (tcgen05 int8) GEMM engine: slicing check, then GEMM vs numpy and vs DMMA."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpc_b200 as G  # noqa: E402
from gpc_b200._lib import check, lib, ptr  # noqa: E402

L = lib()
rng = np.random.default_rng(7)


def slice_check(R, K, kc, S):
    X = rng.standard_normal((R, K)) * np.exp(rng.uniform(-8, 8, (R, 1)))
    Xd = np.ascontiguousarray(X) if kc else np.asfortranarray(X)
    sl = np.zeros((S, R, K), dtype=np.int8)
    sc = np.zeros(R)
    check(L.gpc_oz_slice_check(0, R, K, kc, S, ptr(Xd), ptr(sl), ptr(sc)))
    rec = np.zeros((R, K), dtype=np.longdouble)
    for p in range(S):
        rec += sl[p].astype(np.longdouble) * np.longdouble(2.0) ** (-(8 * p + 6))
    rec *= sc[:, None].astype(np.longdouble)
    err = np.max(np.abs(rec - X.astype(np.longdouble)) / np.max(np.abs(X), axis=1, keepdims=True))
    print("slice R=%d K=%d kc=%d S=%d: max |digit|=%d, rel-to-rowmax err=%.3e (bound %.3e)" % (
        R, K, kc, S, np.abs(sl.astype(int)).max(), float(err), 2.0 ** -(6 + 8 * (S - 1) + 1)), flush=True)


def gemm_case(m, n, k, a_kc, b_kc, lower, cfg, alpha=-1.0, beta=1.0, wide=False):
    A = rng.standard_normal((m, k))
    B = rng.standard_normal((n, k)) if not lower else A
    if wide:
        A = A * np.exp(rng.uniform(-6, 6, (m, 1)))
        B = B * np.exp(rng.uniform(-6, 6, (n, 1))) if not lower else A
    C0 = rng.standard_normal((m, n))
    Ad = np.ascontiguousarray(A) if a_kc else np.asfortranarray(A)
    Bd = np.ascontiguousarray(B) if b_kc else np.asfortranarray(B)
    Cd = np.asfortranarray(C0.copy())
    t0 = time.time()
    check(L.gpc_gemm_check(0, m, n, k, a_kc, b_kc, lower, cfg, alpha, beta, ptr(Ad), ptr(Bd), ptr(Cd)))
    dt = time.time() - t0
    ref = alpha * (A.astype(np.longdouble) @ B.astype(np.longdouble).T) + beta * C0 if m * n * k <= 2 ** 27 else \
        alpha * (A @ B.T) + beta * C0
    scale = np.abs(A) @ np.abs(B).T + np.abs(C0)
    diff = np.abs(Cd - ref) / scale
    if lower:  # only tiles touching the lower triangle are defined
        i, j = np.indices((m, n))
        diff = np.where(j <= i, diff, 0.0)
    print("gemm m=%d n=%d k=%d a_kc=%d b_kc=%d lower=%d cfg=%d wide=%d: max err/(|A||B|) = %.3e  (%.2fs)" % (
        m, n, k, a_kc, b_kc, lower, cfg, wide, float(diff.max()), dt), flush=True)
    return float(diff.max())


if __name__ == "__main__":
    stage = sys.argv[1] if len(sys.argv) > 1 else "all"
    if stage in ("slice", "all"):
        for kc in (0, 1):
            for S in (8, 5):
                slice_check(128, 256, kc, S)
        slice_check(384, 128, 0, 8)
    if stage in ("gemm", "all"):
        gemm_case(128, 128, 128, 0, 0, 0, 0)       # DMMA reference path of the harness
        gemm_case(128, 128, 128, 0, 0, 0, 108)     # one tile, one k-block
        gemm_case(128, 128, 256, 0, 0, 0, 108)
        gemm_case(256, 128, 1024, 0, 0, 0, 108)
        gemm_case(256, 256, 2048, 1, 1, 0, 108)
        gemm_case(256, 256, 512, 0, 1, 0, 108, alpha=1.0, beta=0.0)
        gemm_case(256, 256, 512, 1, 0, 0, 108, wide=True)
        gemm_case(512, 512, 512, 0, 0, 1, 108)
        for S in (7, 6, 5, 4, 3):
            gemm_case(256, 256, 1024, 0, 0, 0, 100 + S)
        gemm_case(1024, 1024, 4096, 0, 0, 0, 108)
        gemm_case(1024, 1024, 4096, 0, 0, 0, 0)
    if stage in ("perf", "all"):
        ms = C.c_double(0)
        for (m, n, k, lower) in [(4096, 4096, 4096, 0), (8192, 8192, 4096, 1), (8192, 8192, 8192, 0), (16384, 16384, 4096, 1)]:
            for cfg in (-1, 108, 107, 106):
                check(L.gpc_bench_gemm(0, m, n, k, 0, 0, lower, cfg, 3, C.byref(ms)))
                fl = (m * (m + 128) * k) if lower else 2.0 * m * n * k
                print("perf m=%d n=%d k=%d lower=%d cfg=%d: %.3f ms  %.1f TFLOP/s-equivalent" % (
                    m, n, k, lower, cfg, ms.value, fl / ms.value / 1e9), flush=True)
