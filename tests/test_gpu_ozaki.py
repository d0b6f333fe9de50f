"""GPU parity of the tensor-core fp64 engine (gpc_b200/csrc/ozaki.cu: tcgen05.mma.kind::i8 + TMEM + TMA) and of the
factor-with-explicit-inverse recursion that feeds it large GEMMs, through the C ABI.  The checker is numpy (long double
where the size allows).  Replaces dgemm_/dsyrk_ below dpotrf_/dpotri_ (reference lapack.h:59-73, 186-222)."""
import os

import numpy as np
import pytest

import gpc_b200 as G
from gpc_b200._lib import check, lib, ptr

pytestmark = pytest.mark.gpu


def _slices(R, K, kc, S, rng):
    X = rng.standard_normal((R, K)) * np.exp(rng.uniform(-10, 10, (R, 1)))
    X[3, :] = 0.0                      # an all-zero row
    X[5, 7] = 0.0
    Xd = np.ascontiguousarray(X) if kc else np.asfortranarray(X)
    sl = np.zeros((S, R, K), dtype=np.int8)
    sc = np.zeros(R)
    check(lib().gpc_oz_slice_check(0, R, K, kc, S, ptr(Xd), ptr(sl), ptr(sc)))
    return X, sl, sc


@pytest.mark.parametrize("kc", [0, 1])
@pytest.mark.parametrize("S,R,K", [(8, 256, 384), (7, 256, 384), (3, 256, 384), (8, 1024, 512)])
def test_slicing_is_exact_to_the_last_digit(kc, S, R, K):
    """x = 2^e_r * sum_p d_p 2^-(8p+6), d_0 in [-65, 65], d_p a balanced byte, remainder below half a unit of the last digit.
    Two shapes: few and many row blocks."""
    rng = np.random.default_rng(11 + S)
    X, sl, sc = _slices(R, K, kc, S, rng)
    assert np.abs(sl[0].astype(int)).max() <= 65
    rec = np.zeros(X.shape, dtype=np.longdouble)
    for p in range(S):
        rec += sl[p].astype(np.longdouble) * np.longdouble(2.0) ** (-(8 * p + 6))
    rec *= sc[:, None].astype(np.longdouble)
    # scales are powers of two strictly above the row maximum (or the smallest normal scale for a zero row)
    m, e = np.frexp(sc)
    assert np.all(m == 0.5)
    rowmax = np.max(np.abs(X), axis=1)
    assert np.all(sc[rowmax > 0] > rowmax[rowmax > 0]) and np.all(sc[rowmax > 0] <= 2 * rowmax[rowmax > 0])
    err = np.abs(rec - X.astype(np.longdouble))
    bound = sc[:, None] * 2.0 ** -(6 + 8 * (S - 1) + 1)
    assert np.all(err <= bound * (1 + 1e-12))
    assert np.all(sl[:, 3, :] == 0)


def _gemm(m, n, k, a_kc, b_kc, flags, cfg, alpha, beta, rng, wide=False, rowmax_metric=False):
    A = rng.standard_normal((m, k))
    B = A if (flags & 1) else rng.standard_normal((n, k))
    if wide:
        A = A * np.exp(rng.uniform(-6, 6, (m, 1)))
        B = A if (flags & 1) else B * np.exp(rng.uniform(-6, 6, (n, 1)))
    a_tri = (flags >> 1) & 3
    b_tri = (flags >> 3) & 3
    ii, kk = np.indices((m, k))
    jj, kj = np.indices((n, k))
    Az = np.where(((a_tri == 1) & (kk < ii)) | ((a_tri == 2) & (kk > ii)), 0.0, A)
    Bz = np.where(((b_tri == 1) & (kj < jj)) | ((b_tri == 2) & (kj > jj)), 0.0, B)
    # the zero part of a triangular operand must never be read: poison it on the device side
    Ap = np.where(Az == A, A, np.nan) if a_tri else A
    Bp = np.where(Bz == B, B, np.nan) if b_tri else B
    if a_tri or b_tri:  # poison only outside the 128-wide diagonal blocks (skipping is tile-granular)
        Ap = np.where((np.abs(kk - ii) < 128) & np.isnan(Ap), 0.0, Ap)
        Bp = np.where((np.abs(kj - jj) < 128) & np.isnan(Bp), 0.0, Bp)
    C0 = rng.standard_normal((m, n))
    Ad = np.ascontiguousarray(Ap) if a_kc else np.asfortranarray(Ap)
    Bd = np.ascontiguousarray(Bp) if b_kc else np.asfortranarray(Bp)
    Cd = np.asfortranarray(C0.copy())
    check(lib().gpc_gemm_check(0, m, n, k, a_kc, b_kc, flags, cfg, alpha, beta, ptr(Ad), ptr(Bd), ptr(Cd)))
    ref = alpha * (Az.astype(np.longdouble) @ Bz.astype(np.longdouble).T) + beta * C0
    if rowmax_metric:
        # the splitting is exact relative to each ROW's largest entry: |dC_ij| <~ sqrt(k) 2^-55 max|a_i| max|b_j|
        scale = (np.max(np.abs(Az), axis=1)[:, None] * np.max(np.abs(Bz), axis=1)[None, :] * np.sqrt(k) * abs(alpha)
                 + np.abs(beta * C0) + 1e-300)
    else:
        scale = abs(alpha) * (np.abs(Az) @ np.abs(Bz).T) + np.abs(beta * C0) + 1e-300
    diff = (np.abs(Cd - ref) / scale).astype(float)
    diff = np.where(np.isnan(Cd), np.inf, diff)
    if flags & 1:  # only the tiles touching the lower triangle are defined
        i, j = np.indices((m, n))
        diff = np.where(j <= i, diff, 0.0)
    return float(diff.max())


@pytest.mark.parametrize("a_kc,b_kc", [(0, 0), (1, 1), (0, 1), (1, 0)])
def test_ozaki_gemm_matches_fp64_all_layouts(a_kc, b_kc):
    rng = np.random.default_rng(100 + 2 * a_kc + b_kc)
    e = _gemm(256, 384, 512, a_kc, b_kc, 0, 108, -1.0, 1.0, rng)
    assert e < 1.5e-16, e   # the final rounding of the fp64 result dominates
    e = _gemm(256, 256, 256, a_kc, b_kc, 0, 108, 0.5, 0.0, rng, wide=True)
    assert e < 1.5e-16, e


def test_ozaki_is_at_least_as_accurate_as_dmma():
    rng = np.random.default_rng(5)
    e_oz = _gemm(512, 512, 2048, 0, 0, 0, 108, -1.0, 1.0, rng)
    rng = np.random.default_rng(5)
    e_dm = _gemm(512, 512, 2048, 0, 0, 0, 0, -1.0, 1.0, rng)
    assert e_oz < 1e-16 and e_oz <= e_dm, (e_oz, e_dm)


@pytest.mark.parametrize("cfg", [108, -1])
@pytest.mark.parametrize("flags", [1, 1 | (1 << 1), 2 << 1, 1 << 3, 2 << 3])
def test_lower_and_triangular_k_ranges(flags, cfg):
    """SYRK (lower tiles), lauum (lower + op(A) zero for kk < i) and the three triangular-operand products of the
    factor-with-inverse recursion; the zero halves of the operands hold NaN and must never be read."""
    rng = np.random.default_rng(40 + flags)
    akc = 1 if flags == (1 | (1 << 1)) else 0
    e = _gemm(512, 512, 512, akc, akc, flags, cfg, -1.0, 1.0, rng, rowmax_metric=True)
    assert e < (1e-15 if cfg >= 100 else 1e-14), (flags, cfg, e)  # the DMMA engine accumulates k = 512 products in fp64


@pytest.mark.parametrize("S,tol", [(7, 2e-15), (6, 5e-13), (5, 1e-10), (3, 5e-6)])
def test_fewer_slices_degrade_as_documented(S, tol):
    rng = np.random.default_rng(3)
    e = _gemm(256, 256, 1024, 0, 0, 0, 100 + S, -1.0, 1.0, rng)
    assert e < tol, (S, e)
    assert e > tol * 1e-4  # and the knob really changes the arithmetic


def test_non_finite_input_poisons_the_row_not_the_matrix():
    rng = np.random.default_rng(9)
    m = n = 128
    k = 256
    A = rng.standard_normal((m, k))
    B = rng.standard_normal((n, k))
    A[17, 5] = np.nan
    C = np.zeros((m, n), order="F")
    check(lib().gpc_gemm_check(0, m, n, k, 0, 0, 0, 108, 1.0, 0.0, ptr(np.asfortranarray(A)), ptr(np.asfortranarray(B)),
                               ptr(C)))
    assert np.all(np.isnan(C[17, :]))
    ok = np.delete(C, 17, axis=0)
    assert np.all(np.isfinite(ok))
    assert np.allclose(ok, np.delete(A, 17, axis=0) @ B.T, rtol=0, atol=1e-12)


def _eval(N, D, mode, ozaki):
    os.environ["GPC_POTRF_MODE"] = mode
    check(lib().gpc_set_gemm_engine(ozaki, 8, 256, 512))
    rng = np.random.default_rng(77)
    X = rng.standard_normal((N, D))
    y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((N, 1))
    kern = G.make_kern(["rbf", "bias", "white"], D)
    kern.setParams([0.3, 1.0, 0.5, 0.02])
    gp = G.CGp(kern, X, y, bias=y.mean(0))
    g, ll = gp.logLikelihoodGradient()
    mu, var = gp.posteriorMeanVar(X[:50] + 0.05)
    gp.ctx.close()
    return ll, g, mu, var


def test_engines_and_recursions_agree_on_a_full_evaluation():
    """N = 4500 (not a multiple of the tile): tensor-core engine + factor-with-inverse (the default) against the DMMA
    engine + recursive TRSM / Schur-complement inverse, same inputs; tolerance 1e-8 relative as for the reference."""
    try:
        ll0, g0, mu0, var0 = _eval(4500, 6, "rec", 0)
        ll1, g1, mu1, var1 = _eval(4500, 6, "winv", 1)
        ll2, g2, _, _ = _eval(4500, 6, "winv", 0)
    finally:
        os.environ.pop("GPC_POTRF_MODE", None)
        check(lib().gpc_set_gemm_engine(1, 8, 256, 512))

    def rel(a, b):
        a, b = np.asarray(a, float), np.asarray(b, float)
        return float(np.max(np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), 1.0)))

    assert rel(ll1, ll0) < 1e-9 and rel(g1, g0) < 1e-8, (rel(ll1, ll0), rel(g1, g0))
    assert rel(ll2, ll0) < 1e-9 and rel(g2, g0) < 1e-8
    assert rel(mu1, mu0) < 1e-8 and rel(var1, var0) < 1e-8


def test_posterior_after_factorisation_only_uses_block_substitution():
    """gpc_potrf -> gpc_solve_alpha -> gpc_posterior without gpc_inverse: the top-level W21 is still deferred, so the
    variance solve V = K* L^-T runs by block substitution with the diagonal halves of W (CGp::_posteriorVar,
    CGp.cpp:600-613); the same call after gpc_inverse uses the single triangular product.  Both must agree with the
    oracle."""
    import ctypes as C
    from oracle import gp_oracle as O
    rng = np.random.default_rng(31)
    N, D, Ns = 1500, 4, 300
    X = rng.standard_normal((N, D))
    y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((N, 1))
    Xs = rng.standard_normal((Ns, D))
    types = ["rbf", "white"]
    tp = np.array([-0.7, 0.2, -2.0])
    kern = G.make_kern(types, D, tp)
    mu_ref, var_ref = O.gp_posterior(O.kern_from_trans(types, tp, D), X, y, Xs)
    ctx = G.DeviceContext(N, D, 1)
    try:
        ctx.set_X(X)
        ctx.set_M(y)
        arr, n, keep = kern._kcomps()
        L = lib()
        check(L.gpc_kern_build(ctx.handle, arr, n))
        info, logdet, quad = C.c_int(0), C.c_double(0), C.c_double(0)
        assert check(L.gpc_potrf(ctx.handle, C.byref(info), C.byref(logdet))) == 0
        check(L.gpc_solve_alpha(ctx.handle, C.byref(quad)))
        Xf = np.asfortranarray(Xs)
        for with_inverse in (False, True):
            if with_inverse:
                check(L.gpc_inverse(ctx.handle))
            mu = np.zeros((Ns, 1), order="F")
            var = np.zeros((Ns, 1), order="F")
            check(L.gpc_posterior(ctx.handle, arr, n, ptr(Xf), Ns, Ns, ptr(mu), ptr(var)))
            assert np.max(np.abs(mu - mu_ref) / np.maximum(1.0, np.abs(mu_ref))) < 1e-8, with_inverse
            assert np.max(np.abs(var - var_ref) / np.maximum(1.0, np.abs(var_ref))) < 1e-8, with_inverse
    finally:
        ctx.close()
