import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_ozaki as T
import gpc_b200 as G
from gpc_b200._lib import check, lib, ptr
import numpy as _np
orig = T.check
def run(flags, cfg, akc):
    rng = np.random.default_rng(40 + flags)
    m = n = k = 512
    A = rng.standard_normal((m, k)); B = A if (flags & 1) else rng.standard_normal((n, k))
    a_tri = (flags >> 1) & 3; b_tri = (flags >> 3) & 3
    ii, kk = np.indices((m, k)); jj, kj = np.indices((n, k))
    Az = np.where(((a_tri == 1) & (kk < ii)) | ((a_tri == 2) & (kk > ii)), 0.0, A)
    Bz = np.where(((b_tri == 1) & (kj < jj)) | ((b_tri == 2) & (kj > jj)), 0.0, B)
    Ap = np.where(Az == A, A, np.nan) if a_tri else A
    Bp = np.where(Bz == B, B, np.nan) if b_tri else B
    if a_tri or b_tri:
        Ap = np.where((np.abs(kk - ii) < 128) & np.isnan(Ap), 0.0, Ap)
        Bp = np.where((np.abs(kj - jj) < 128) & np.isnan(Bp), 0.0, Bp)
    C0 = rng.standard_normal((m, n))
    Ad = np.ascontiguousarray(Ap) if akc else np.asfortranarray(Ap)
    Bd = np.ascontiguousarray(Bp) if akc else np.asfortranarray(Bp)
    Cd = np.asfortranarray(C0.copy())
    check(lib().gpc_gemm_check(0, m, n, k, akc, akc, flags, cfg, -1.0, 1.0, ptr(Ad), ptr(Bd), ptr(Cd)))
    ref = -(Az.astype(np.longdouble) @ Bz.astype(np.longdouble).T) + C0
    i, j = np.indices((m, n))
    low = (j <= i) if (flags & 1) else np.ones((m, n), bool)
    bad = np.isnan(Cd) & low
    rs = np.max(np.abs(Az), axis=1)[:, None] * np.max(np.abs(Bz), axis=1)[None, :] * np.sqrt(k) + np.abs(C0)
    err = np.where(low & ~np.isnan(Cd), np.abs(Cd - ref).astype(float), 0)
    print("flags", flags, "cfg", cfg, "nan", bad.sum(), "rows", np.unique(i[bad])[:6], "cols", np.unique(j[bad])[:6],
          "max abs err %.2e  max err/(rowmax*rowmax*sqrt(k)+|C|) %.2e" % (err.max(), (err / rs).max()), "nan in Ap read region?", flush=True)
for flags, akc in ((3, 1), (3, 0), (4, 0), (8, 0), (16, 0), (1, 0)):
    for cfg in (108, -1):
        run(flags, cfg, akc)
