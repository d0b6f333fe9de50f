"""tools/gpu_diag.py -- one-shot GPU bring-up diagnostics (prints errors instead of asserting) so that a single
gpurun round trip localises a bug to a kernel.  Not part of the test suite."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpc_b200 as G  # noqa: E402
from gpc_b200 import matrix as M  # noqa: E402
from oracle import gp_oracle as O  # noqa: E402

rng = np.random.default_rng(0)


def section(name):
    print("\n==== " + name, flush=True)


def guarded(fn):
    try:
        fn()
    except Exception:
        traceback.print_exc()
        sys.stdout.flush()


def t_gemm():
    section("dgemm / dsyrk")
    for (m, n, k) in [(11, 22, 11), (130, 70, 50), (256, 384, 300), (1000, 900, 777), (2500, 2300, 640)]:
        for ta in "nt":
            for tb in "nt":
                A = rng.standard_normal((m, k) if ta == "n" else (k, m))
                B = rng.standard_normal((k, n) if tb == "n" else (n, k))
                C0 = rng.standard_normal((m, n))
                ref = 0.7 * (A if ta == "n" else A.T) @ (B if tb == "n" else B.T) - 0.3 * C0
                got = M.gemm(C0, A, B, 0.7, -0.3, ta, tb)
                print(f"gemm {m}x{n}x{k} {ta}{tb} maxerr {np.abs(got - ref).max():.2e}", flush=True)
    for (n, k) in [(11, 16), (300, 129), (1500, 700)]:
        for tr in "nt":
            A = rng.standard_normal((n, k) if tr == "n" else (k, n))
            C0 = rng.standard_normal((n, n))
            C0 = C0 + C0.T
            full = 1.3 * (A @ A.T if tr == "n" else A.T @ A) + 0.5 * C0
            for ul in "ul":
                got = M.syrk(C0, A, 1.3, 0.5, ul, tr)
                tri = np.triu if ul == "u" else np.tril
                other = np.tril if ul == "u" else np.triu
                e1 = np.abs(tri(got) - tri(full)).max()
                e2 = np.abs(other(got, -1 if ul == "u" else 1) - other(C0, -1 if ul == "u" else 1)).max()
                print(f"syrk n{n} k{k} {ul}{tr} err {e1:.2e} untouched {e2:.2e}", flush=True)


def spd(n, cond=1e3):
    B = rng.standard_normal((n, n))
    Q, _ = np.linalg.qr(B)
    ev = np.logspace(0, np.log10(cond), n)
    return (Q * ev) @ Q.T


def t_potrf():
    section("dpotrf / dpotri / dtrsm")
    for n in [11, 128, 129, 200, 500, 1000, 2048, 3000]:
        A = spd(n)
        t0 = time.time()
        U = M.chol(A, "U")
        dt = time.time() - t0
        Lr = np.linalg.cholesky(A)
        eU = np.abs(U - Lr.T).max()
        Lm = M.chol(A, "L")
        eL = np.abs(Lm - Lr).max()
        t0 = time.time()
        inv = M.pdinv(U)
        dt2 = time.time() - t0
        eI = np.abs(inv - np.linalg.inv(A)).max() / np.abs(np.linalg.inv(A)).max()
        invL = M.potri(Lm, "L")
        eIL = np.abs(np.tril(invL) - np.tril(np.linalg.inv(A))).max() / np.abs(np.linalg.inv(A)).max()
        print(f"n={n}: chol U err {eU:.2e} L err {eL:.2e} ({dt*1e3:.1f} ms) | pdinv rel {eI:.2e} potri L rel {eIL:.2e} ({dt2*1e3:.1f} ms)", flush=True)
    A = np.eye(300)
    A[170, 170] = -1.0
    try:
        M.potrf(A, "L")
        print("non-PD: NO EXCEPTION (bad)")
    except G.MatrixNonPosDef as e:
        print("non-PD info", e.info, "(expect 171)")
    for (m, n) in [(16, 30), (200, 150), (700, 513)]:
        B = rng.standard_normal((m, n))
        for side in "lr":
            k = m if side == "l" else n
            T0 = rng.standard_normal((k, k)) + 4 * np.eye(k) * np.sqrt(k)
            for ul in "ul":
                T = np.triu(T0) if ul == "u" else np.tril(T0)
                for tr in "nt":
                    for dg in "nu":
                        Tt = T.copy()
                        if dg == "u":
                            np.fill_diagonal(Tt, 1.0)
                        op = Tt if tr == "n" else Tt.T
                        ref = 0.9 * (np.linalg.solve(op, B) if side == "l" else np.linalg.solve(op.T, B.T).T)
                        got = M.trsm(B, T, 0.9, side, ul, tr, dg)
                        err = np.abs(got - ref).max() / max(1.0, np.abs(ref).max())
                        print(f"trsm {m}x{n} {side}{ul}{tr}{dg} relerr {err:.2e}", flush=True)
    A = spd(400)
    x = rng.standard_normal(400)
    y0 = rng.standard_normal(400)
    for ul in "ul":
        got = M.symv(y0, A, x, 1.5, 0.25, ul)
        print(f"symv {ul} err {np.abs(got - (1.5 * A @ x + 0.25 * y0)).max():.2e}")


def t_kern():
    section("kernels vs numpy oracle")
    f = np.load(os.path.join(ROOT, "tests/golden/kern_matfiles.npz"))
    for name in ["rbf", "rbfard", "matern32", "matern52", "lin", "poly", "white", "bias"]:
        X, X2, tp = f[name + "_X"], f[name + "_X2"], f[name + "_params"]
        kern = G.make_kern([name], X.shape[1], tp)
        K = kern.compute(X)
        K4 = kern.compute(X, X2)
        kd = kern.diagCompute(X)
        g2 = kern.getGradTransParams(X, f[name + "_covGrad"])
        print(f"{name:9s} K {np.abs(K - f[name+'_K2']).max():.1e} K4 {np.abs(K4 - f[name+'_K4']).max():.1e} "
              f"diag {np.abs(kd - f[name+'_k2']).max():.1e} g2 {np.abs(g2 - f[name+'_g2']).max():.1e}", flush=True)


def t_gp():
    section("CGp vs golden reference outputs")
    f = np.load(os.path.join(ROOT, "tests/golden/random_reference.npz"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import CASES, rel_err
    for tag, types in CASES.items():
        X, y = f[tag + "_X"], f[tag + "_y"]
        kern = G.make_kern(types, X.shape[1], f[tag + "_tparams"])
        K = kern.compute(X)
        Kx = kern.compute(X, f[tag + "_X2"])
        g = kern.getGradTransParams(X, f[tag + "_covGrad"])
        gp = G.CGp(kern, X, y, bias=f[tag + "_bias"], scale=f[tag + "_scale"])
        gl, ll = gp.logLikelihoodGradient()
        mu, var = gp.posteriorMeanVar(f[tag + "_X2"])
        print(f"{tag:18s} K {np.abs(K - f[tag+'_K']).max():.1e} Kx {np.abs(Kx - f[tag+'_Kx']).max():.1e} "
              f"g {rel_err(g, f[tag+'_g']):.1e} | ll {rel_err(ll, f[tag+'_ll']):.1e} gll {rel_err(gl, f[tag+'_gll']):.1e} "
              f"mu {rel_err(mu, f[tag+'_mu']):.1e} var {rel_err(var, f[tag+'_var']):.1e}", flush=True)
    g = np.load(os.path.join(ROOT, "tests/golden/gp_reference.npz"))
    kern = G.make_kern(["rbf", "bias", "white"], 1, g["sinc_params"])
    gp = G.CGp(kern, g["sinc_X"], g["sinc_y"], bias=g["sinc_bias"])
    gl, ll = gp.logLikelihoodGradient()
    print("sinc ll", ll, "expect -28.2080301154265 ; g", gl, "expect", g["sinc_g_ref"])
    kern = G.make_kern(["rbf", "lin", "bias", "white"], 2, g["ftc_params"])
    gp = G.CGp(kern, g["ftc_X"], g["ftc_y"], bias=g["ftc_bias"])
    gl, ll = gp.logLikelihoodGradient()
    print("ftc ll", ll, "expect", float(g["ftc_ll_ref"]), "; g err", np.abs(gl - g["ftc_grads_matlab"]).max())
    # GP-LVM
    kern = G.make_kern(["rbf", "bias", "white"], 2, f["lvm_tparams"])
    lvm = G.CGplvm(kern, f["lvm_m"], f["lvm_X"])
    gl, ll = lvm.logLikelihoodGradient()
    print(f"gplvm ll {rel_err(ll, float(f['lvm_ll'])):.1e} g {rel_err(gl, f['lvm_g']):.1e} (kern part {rel_err(gl[:4], f['lvm_g'][:4]):.1e})")
    kern = G.make_kern(["rbf", "lin", "matern32", "white"], 2, f["lvm2_tparams"])
    lvm = G.CGplvm(kern, f["lvm_m"], f["lvm2_X"])
    gl, ll = lvm.logLikelihoodGradient()
    print(f"gplvm2 ll {rel_err(ll, float(f['lvm2_ll'])):.1e} g {rel_err(gl, f['lvm2_g']):.1e} (kern part {rel_err(gl[:6], f['lvm2_g'][:6]):.1e})")


def t_big():
    section("performance probes")
    import ctypes as C
    tf = C.c_double(0)
    G._lib.check(G.lib().gpc_bench_dmma_peak(0, C.byref(tf)))
    print(f"DMMA register-resident peak: {tf.value:.2f} TFLOP/s")
    for (n, k) in [(4096, 4096), (8192, 2048), (8192, 8192), (16384, 4096)]:
        ms = C.c_double(0)
        G._lib.check(G.lib().gpc_bench_syrk(0, n, k, 3, C.byref(ms)))
        fl = n * (n + 128) * k  # lower tiles incl. diagonal
        print(f"syrk n={n} k={k}: {ms.value:.3f} ms  {fl / ms.value / 1e9:.2f} TFLOP/s")
    for (N, D, types) in [(2048, 8, ["rbf", "white"]), (8192, 8, ["rbf", "white"]), (16384, 16, ["rbfard", "white"])]:
        rg = np.random.default_rng(20261017)
        X = rg.standard_normal((N, D))
        y = np.sin(X[:, :1]) + 0.1 * rg.standard_normal((N, 1))
        y = y - y.mean()
        if types[0] == "rbf":
            kern = G.make_kern(types, D)
            kern.setParams([1.0 / D, 1.0, 0.01])
        else:
            kern = G.make_kern(types, D)
            kern.setParams(np.concatenate([[1.0 / D, 1.0], 0.25 + 0.5 * np.arange(D) / (D - 1), [0.01]]))
        gp = G.CGp(kern, X, y)
        for rep in range(3):
            gp.KupToDate = False
            t0 = time.time()
            g, ll = gp.logLikelihoodGradient()
            dt = time.time() - t0
        print(f"N={N} D={D} {types}: eval wall {dt*1e3:.1f} ms ll {ll:.6f} |g| {np.abs(g).max():.3e} launches {gp.ctx.launch_count()} phases {gp.timings()}", flush=True)
        if N <= 8192:
            r = O.gp_loglik_grad(O.kern_from_trans(types, kern.getTransParams(), D), X, y) if N <= 2048 else None
            if r:
                print("   vs numpy oracle: ll rel", abs(ll - r["ll"]) / abs(r["ll"]), "g rel", np.abs(g - r["g"]).max() / np.abs(r["g"]).max())


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "potrf", "kern", "gp", "big"]
    print("devices:", G.lib().gpc_device_count())
    for w in which:
        guarded({"gemm": t_gemm, "potrf": t_potrf, "kern": t_kern, "gp": t_gp, "big": t_big}[w])
