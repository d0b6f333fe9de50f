"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ctypes) behind the mirrored CKern / CGp /
CMatrix interface, against (a) the reference's MATLAB golden vectors, (b) outputs of the unmodified reference on
seeded inputs (tests/golden), (c) the numpy oracle on fresh random inputs, and (d) size-independent properties at
BASELINE.json's full sizes.  Tolerances: K entries 1e-12 abs; ll / gradients / posterior 1e-8 relative
(north_star); dense primitives the reference's own MATCHTOL 1e-10 / 1e-8 (ndlutil.h:33, testMatrix.cpp:608)."""
import numpy as np
import pytest
import scipy.linalg as sla

import gpc_b200 as G
from gpc_b200 import matrix as M
from conftest import CASES, SINGLE, rel_err
from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu
MATCHTOL = 1e-10


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert G.lib().gpc_device_count() > 0, "no CUDA device: the gpu tests must run on the B200 box"


# ---------------------------------------------------------------------------------------------------------
# kernels (testKern.cpp:236-376)
@pytest.mark.parametrize("name", SINGLE)
def test_kernel_matlab_fixtures(kern_mat, name):
    f = kern_mat
    X, X2 = f[name + "_X"], f[name + "_X2"]
    kern = G.make_kern([name], X.shape[1], f[name + "_params"])
    assert np.abs(kern.compute(X) - f[name + "_K2"]).max() < 1e-12 * max(1.0, np.abs(f[name + "_K2"]).max())
    assert np.abs(kern.compute(X, X2) - f[name + "_K4"]).max() < 1e-12 * max(1.0, np.abs(f[name + "_K4"]).max())
    assert np.abs(kern.diagCompute(X) - f[name + "_k2"]).max() < 1e-12 * max(1.0, np.abs(f[name + "_k2"]).max())
    assert np.abs(kern.getGradTransParams(X, f[name + "_covGrad"]) - f[name + "_g2"]).max() < MATCHTOL


@pytest.mark.parametrize("tag", list(CASES))
def test_compound_kernel_vs_reference(rand_ref, tag):
    f, types = rand_ref, CASES[tag]
    X, X2 = f[tag + "_X"], f[tag + "_X2"]
    kern = G.make_kern(types, X.shape[1], f[tag + "_tparams"])
    np.testing.assert_allclose(kern.params, f[tag + "_params"], rtol=1e-14)
    K = kern.compute(X)
    assert np.abs(K - f[tag + "_K"]).max() < 1e-12 * max(1.0, np.abs(f[tag + "_K"]).max())
    assert np.array_equal(K, K.T)
    assert np.abs(kern.compute(X, X2) - f[tag + "_Kx"]).max() < 1e-12 * max(1.0, np.abs(f[tag + "_Kx"]).max())
    assert np.abs(kern.diagCompute(X2) - f[tag + "_kdiag"]).max() < 1e-12 * max(1.0, np.abs(f[tag + "_kdiag"]).max())
    assert rel_err(kern.getGradTransParams(X, f[tag + "_covGrad"]), f[tag + "_g"]) < 1e-10


def test_kernel_edge_shapes():
    """ragged sizes around the 64/128 tile edges, N=1, duplicate rows (r = 0 off the diagonal), D=1 and D=40."""
    rng = np.random.default_rng(5)
    for N, D in [(1, 1), (2, 3), (63, 2), (64, 1), (65, 5), (127, 40), (128, 7), (129, 4), (257, 3)]:
        X = rng.standard_normal((N, D))
        if N > 2:
            X[1] = X[0]
        types = ["rbf", "matern32", "matern52", "rbfard", "lin", "poly", "bias", "white"]
        tp = 0.3 * rng.standard_normal(sum(O.nparams(t, D) for t in types))
        kern = G.make_kern(types, D, tp)
        ref = O.kern_compute(O.kern_from_trans(types, tp, D), X)
        K = kern.compute(X)
        assert K.shape == (N, N)
        assert np.abs(K - ref).max() < 1e-11 * max(1.0, np.abs(ref).max()), (N, D)
        assert not np.isnan(K).any()


# ---------------------------------------------------------------------------------------------------------
# dense primitives (testMatrix.cpp)
def test_cholesky_inverse_matlab(matrix_mat):
    f = matrix_mat
    C_ = f["choleskyMatrixTest_C"]
    assert np.abs(M.chol(C_, "U") - f["choleskyMatrixTest_U"]).max() < MATCHTOL   # testCholesky :206-236
    assert np.abs(M.chol(C_, "L") - f["choleskyMatrixTest_L"]).max() < MATCHTOL
    A = f["invMatrixTest_A"]
    try:
        U = M.chol(A, "U")
    except G.MatrixNonPosDef:
        U = None
    if U is not None:
        assert np.abs(M.pdinv(U) - f["invMatrixTest_Ainv"]).max() < 1e-8              # testInv :187-205


def test_syrk_gemm_matlab(matrix_mat):
    f = matrix_mat
    a, b = f["syrkMatrixTest_alpha"].item(), f["syrkMatrixTest_beta"].item()
    A, Cm, Dm = f["syrkMatrixTest_A"], f["syrkMatrixTest_C"], f["syrkMatrixTest_D"]
    for ul in "ul":   # testSyrk :332-393 (un/ln/ut/lt)
        tri = np.triu if ul == "u" else np.tril
        assert np.abs(tri(M.syrk(Cm, A, a, b, ul, "n")) - tri(f["syrkMatrixTest_SYRK1"])).max() < MATCHTOL
        assert np.abs(tri(M.syrk(Dm, A, a, b, ul, "t")) - tri(f["syrkMatrixTest_SYRK2"])).max() < MATCHTOL
    a, b = f["gemmMatrixTest_alpha"].item(), f["gemmMatrixTest_beta"].item()
    D_, E_, F_, G_, H_ = (f["gemmMatrixTest_" + k] for k in "DEFGH")
    G1, G2 = f["gemmMatrixTest_GEMM1"], f["gemmMatrixTest_GEMM2"]
    # testGemm :266-331: F.gemm(D,E,nn); G.gemm(D,E,tt); GEMM1.gemm(D,H,nt); GEMM2.gemm(D,H,tn)
    assert np.abs(M.gemm(F_, D_, E_, a, b, "n", "n") - G1).max() < MATCHTOL
    assert np.abs(M.gemm(G_, D_, E_, a, b, "t", "t") - G2).max() < MATCHTOL
    assert np.abs(M.gemm(G1, D_, H_, a, b, "n", "t") - f["gemmMatrixTest_GEMM3"]).max() < MATCHTOL
    assert np.abs(M.gemm(G2, D_, H_, a, b, "t", "n") - f["gemmMatrixTest_GEMM4"]).max() < MATCHTOL


def test_trsm_all_sixteen_variants():
    """testTrsm (testMatrix.cpp:606-836, tol 1e-8): side x uplo x trans x diag, at sizes that cross tile edges."""
    rng = np.random.default_rng(11)
    for (m, n) in [(16, 30), (200, 150), (300, 513)]:
        B = rng.standard_normal((m, n))
        for side in "lr":
            k = m if side == "l" else n
            T0 = rng.standard_normal((k, k)) / np.sqrt(k) + 2.0 * np.eye(k)
            for ul in "ul":
                T = np.triu(T0) if ul == "u" else np.tril(T0)
                for tr in "nt":
                    for dg in "nu":
                        got = M.trsm(B, T, 0.9, side, ul, tr, dg)
                        rhs = B if side == "l" else B.T
                        # right side: X op(T) = aB  <=>  op(T)' X' = aB'
                        t_eff = (tr == "t") if side == "l" else (tr == "n")
                        ref = 0.9 * sla.solve_triangular(T, rhs, lower=(ul == "l"), trans=1 if t_eff else 0,
                                                         unit_diagonal=(dg == "u"))
                        ref = ref if side == "l" else ref.T
                        assert rel_err(got, ref) < 1e-8, (m, n, side, ul, tr, dg)


def test_trsm_matlab_fixture_all_sixteen(matrix_mat):
    """testTrsm (testMatrix.cpp:606-836): the reference's own known answers TRSM1..16 of matfiles/trsmMatrixTest.mat, in its
    order -- (L | L2 | U | U2) x (left | right) x (N | T) x (non-unit | unit) -- at its tolerance 1e-8, taken relative to the
    largest expected entry: the unit-diagonal and U2 cases have solutions of size 1e12 (the fixture's triangular factors are
    random, not well conditioned), where 1e-8 absolute is below one ulp."""
    f = matrix_mat
    B, alpha = f["trsmMatrixTest_B"], float(np.ravel(f["trsmMatrixTest_alpha"])[0])
    order = [("L", "l", "l", "n", "n"), ("L", "l", "l", "t", "n"), ("L2", "r", "l", "n", "n"), ("L2", "r", "l", "t", "n"),
             ("L", "l", "l", "n", "u"), ("L", "l", "l", "t", "u"), ("L2", "r", "l", "n", "u"), ("L2", "r", "l", "t", "u"),
             ("U", "l", "u", "n", "n"), ("U", "l", "u", "t", "n"), ("U2", "r", "u", "n", "n"), ("U2", "r", "u", "t", "n"),
             ("U", "l", "u", "n", "u"), ("U", "l", "u", "t", "u"), ("U2", "r", "u", "n", "u"), ("U2", "r", "u", "t", "u")]
    for k, (mat, side, ul, tr, dg) in enumerate(order, start=1):
        got = M.trsm(B, f["trsmMatrixTest_" + mat], alpha, side, ul, tr, dg)
        want = f["trsmMatrixTest_TRSM%d" % k]
        assert np.abs(got - want).max() < 1e-8 * max(1.0, np.abs(want).max()), (k, mat, side, ul, tr, dg)


def test_syr_matlab_fixture(matrix_mat):
    """testSyr (testMatrix.cpp:1138-1203): A + alpha x x' on one triangle, mirrored -- SYR1 (x a column), SYR2 (x = row i of
    B: stride = the leading dimension, CMatrix::syrRow CMatrix.h:535-542), SYR3 (x = column j of B), both triangles; and a
    size across a tile edge against numpy."""
    import ctypes as C
    from gpc_b200._lib import check, lib, ptr
    f = matrix_mat
    A, B, x = f["syrMatrixTest_A"], np.asfortranarray(f["syrMatrixTest_B"]), np.ravel(f["syrMatrixTest_x"])
    alpha = float(np.ravel(f["syrMatrixTest_alpha"])[0])
    i, j = int(np.ravel(f["syrMatrixTest_i"])[0]) - 1, int(np.ravel(f["syrMatrixTest_j"])[0]) - 1
    for ul in "ul":
        assert np.abs(M.syr(A, x, alpha, ul) - f["syrMatrixTest_SYR1"]).max() < MATCHTOL
        assert np.abs(M.syr(A, B[:, j], alpha, ul) - f["syrMatrixTest_SYR3"]).max() < MATCHTOL
        # row i of B through the stride argument (dsyr_'s incx), as syrRow passes it
        F = np.asfortranarray(A.copy())
        n = F.shape[0]
        check(lib().gpc_dsyr(0, ul.encode(), n, alpha, C.c_void_p(B.ctypes.data + 8 * i), B.shape[0], ptr(F), n))
        tri = np.triu(F) if ul == "u" else np.tril(F)
        assert np.abs(tri + tri.T - np.diag(np.diag(F)) - f["syrMatrixTest_SYR2"]).max() < MATCHTOL
    rng = np.random.default_rng(2)
    n = 300
    S0 = rng.standard_normal((n, n))
    S0 = S0 + S0.T
    v = rng.standard_normal(n)
    assert np.abs(M.syr(S0, v, -0.7, "l") - (S0 - 0.7 * np.outer(v, v))).max() < 1e-12


def test_potrf_sizes_and_nonpd():
    rng = np.random.default_rng(3)
    for n in [1, 5, 127, 128, 129, 300, 1000, 2050]:
        B = rng.standard_normal((n, n))
        A = B @ B.T / n + np.eye(n)
        L = M.chol(A, "L")
        assert np.abs(L @ L.T - A).max() < 1e-11 * n
        assert np.abs(L - np.linalg.cholesky(A)).max() < 1e-10
        Ainv = M.pdinv(M.chol(A, "U"))
        assert np.abs(Ainv @ A - np.eye(n)).max() < 1e-9
        assert np.array_equal(Ainv, Ainv.T)
    A = np.eye(300)
    A[170, 170] = -1.0
    with pytest.raises(G.MatrixNonPosDef) as ei:       # info = order of the first bad pivot (dpotrf_ semantics)
        M.potrf(A, "L")
    assert ei.value.info == 171
    A = np.eye(40)
    A[0, 0] = np.nan
    with pytest.raises(G.MatrixNonPosDef):
        M.potrf(A, "U")


def test_jitchol_vs_reference(rand_ref):
    f = rand_ref
    U, jit, Aout = M.jitChol(f["jit_A"])
    assert rel_err(jit, float(f["jit_val"])) < 1e-12
    assert np.abs(Aout - f["jit_Aout"]).max() < 1e-10
    assert np.abs(U.T @ U - Aout).max() < 1e-8


def test_symv():
    rng = np.random.default_rng(4)
    A = rng.standard_normal((400, 400))
    A = A + A.T
    x, y0 = rng.standard_normal(400), rng.standard_normal(400)
    for ul in "ul":
        junk = np.triu(A) if ul == "u" else np.tril(A)   # only the named triangle may be referenced
        assert np.abs(M.symv(y0, junk, x, 1.5, 0.25, ul) - (1.5 * A @ x + 0.25 * y0)).max() < 1e-10


# ---------------------------------------------------------------------------------------------------------
# CGp (testGp.cpp:118-150 + reference outputs)
@pytest.mark.parametrize("tag", list(CASES))
def test_gp_eval_vs_reference(rand_ref, tag):
    f, types = rand_ref, CASES[tag]
    X, y = f[tag + "_X"], f[tag + "_y"]
    kern = G.make_kern(types, X.shape[1], f[tag + "_tparams"])
    gp = G.CGp(kern, X, y, bias=f[tag + "_bias"], scale=f[tag + "_scale"])
    g, ll = gp.logLikelihoodGradient()
    assert rel_err(ll, float(f[tag + "_ll"])) < 1e-8
    assert rel_err(g, f[tag + "_gll"]) < 1e-8
    mu, var = gp.posteriorMeanVar(f[tag + "_X2"])
    assert rel_err(mu, f[tag + "_mu"]) < 1e-8
    assert rel_err(var, f[tag + "_var"]) < 1e-8
    # device-resident matrices as the reference exposes them under -DDBG (CGp.h:359-361)
    K = gp.ctx.download(0)
    L = gp.ctx.download(1)
    Kinv = gp.ctx.download(2)
    assert np.abs(L @ L.T - K).max() < 1e-10 * np.abs(K).max()
    assert np.abs(Kinv @ K - np.eye(K.shape[0])).max() < 1e-7
    assert np.all(np.triu(L, 1) == 0.0)


def test_gp_ftc_matlab_and_sinc(gp_ref):
    f = gp_ref
    kern = G.make_kern(["rbf", "lin", "bias", "white"], 2, f["ftc_params"])
    gp = G.CGp(kern, f["ftc_X"], f["ftc_y"], bias=f["ftc_bias"])
    g, ll = gp.logLikelihoodGradient()
    N = f["ftc_X"].shape[0]
    assert abs(ll + N * O.HALFLOGTWOPI - float(f["ftc_ll_matlab"])) < 1e-8      # MATLAB ll omits -N/2 log 2pi
    assert rel_err(g, f["ftc_grads_matlab"]) < 1e-8
    mu, var = gp.posteriorMeanVar(f["ftc_Xs"])
    assert rel_err(mu, f["ftc_mu_ref"]) < 1e-8 and rel_err(var, f["ftc_var_ref"]) < 1e-8
    # config 1: examples/sinc.svml, gp learn defaults
    kern = G.make_kern(["rbf", "bias", "white"], 1, f["sinc_params"])
    gp = G.CGp(kern, f["sinc_X"], f["sinc_y"], bias=f["sinc_bias"])
    g, ll = gp.logLikelihoodGradient()
    assert abs(ll - (-28.2080301154265)) < 1e-8 * 28.2
    assert rel_err(g, [-4.6975707918852, -10.9705430321513, -0.33705464079873, -7.92625419624426]) < 1e-8
    mu, var = gp.posteriorMeanVar(f["sinc_Xs"])
    assert rel_err(mu, f["sinc_mu_ref"]) < 1e-8 and rel_err(var, f["sinc_var_ref"]) < 1e-8


def test_gp_jitter_retry_matches_oracle():
    """duplicate inputs + tiny noise: the first factorisation fails and the jitChol schedule kicks in."""
    rng = np.random.default_rng(9)
    X = rng.standard_normal((150, 2))
    X[100:] = X[:50]
    y = rng.standard_normal((150, 1))
    types = ["rbf", "white"]
    tp = np.array([0.0, 0.0, -40.0])
    r = O.gp_loglik_grad(O.kern_from_trans(types, tp, 2), X, y)
    gp = G.CGp(G.make_kern(types, 2, tp), X, y)
    g, ll = gp.logLikelihoodGradient()
    assert gp._out[2] > 0.0
    assert rel_err(ll, r["ll"]) < 1e-5      # ill-conditioned by construction (cond ~ 1e8): looser tolerance
    assert np.isfinite(g).all()


def test_gplvm_vs_reference(rand_ref):
    f = rand_ref
    lvm = G.CGplvm(G.make_kern(["rbf", "bias", "white"], 2, f["lvm_tparams"]), f["lvm_m"], f["lvm_X"])
    g, ll = lvm.logLikelihoodGradient()
    assert rel_err(ll, float(f["lvm_ll"])) < 1e-8
    assert rel_err(g, f["lvm_g"]) < 1e-8
    lvm = G.CGplvm(G.make_kern(["rbf", "lin", "matern32", "white"], 2, f["lvm2_tparams"]), f["lvm_m"], f["lvm2_X"])
    g, ll = lvm.logLikelihoodGradient()
    assert rel_err(ll, float(f["lvm2_ll"])) < 1e-8
    assert rel_err(g, f["lvm2_g"]) < 1e-8


@pytest.mark.parametrize("types,N,D,d", [(["rbf", "white"], 700, 8, 1), (["rbfard", "white"], 513, 16, 1),
                                          (["matern52", "white"], 640, 32, 2), (["poly", "lin", "bias", "white"], 384, 6, 1)])
def test_gp_random_vs_numpy_oracle(types, N, D, d):
    rng = np.random.default_rng(N + D)
    X = rng.standard_normal((N, D))
    y = np.sin(X[:, :1]) @ np.ones((1, d)) + 0.1 * rng.standard_normal((N, d))
    kern = G.make_kern(types, D)
    tp = kern.getTransParams()
    tp[-1] = np.log(0.01)
    if types[0] in ("rbf", "rbfard"):
        tp[0] = np.log(1.0 / D)
    if types[0] == "matern52":
        tp[0] = np.log(np.sqrt(D))
    if types[0] == "poly":
        tp[:4] = np.log([0.05, 0.5, 0.3, 0.1])
    kern.setTransParams(tp)
    r = O.gp_loglik_grad(O.kern_from_trans(types, tp, D), X, y)
    gp = G.CGp(kern, X, y)
    g, ll = gp.logLikelihoodGradient()
    assert rel_err(ll, r["ll"]) < 1e-8
    assert rel_err(g, r["g"]) < 1e-8
    # finite-difference pin (COptimisable::checkGradients, COptimisable.cpp:9-44, h = 1e-6)
    for i in range(min(3, tp.size)):
        tpp, tpm = tp.copy(), tp.copy()
        tpp[i] += 1e-6
        tpm[i] -= 1e-6
        gp.setOptParams(tpp)
        lp = gp.logLikelihood()
        gp.setOptParams(tpm)
        lm = gp.logLikelihood()
        assert abs((lp - lm) / 2e-6 - g[i]) < 1e-4 * max(1.0, abs(g[i]))


# ---------------------------------------------------------------------------------------------------------
# full BASELINE sizes: size-independent properties (the oracle cannot run these in seconds)
def _config(N, D, kind):
    rng = np.random.default_rng(20261017)
    X = rng.standard_normal((N, D))
    y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((N, 1))
    y = y - y.mean()
    if kind == "rbf":
        kern = G.make_kern(["rbf", "white"], D)
        kern.setParams([1.0 / D, 1.0, 0.01])
    elif kind == "rbfard":
        kern = G.make_kern(["rbfard", "white"], D)
        kern.setParams(np.concatenate([[1.0 / D, 1.0], 0.25 + 0.5 * np.arange(D) / (D - 1), [0.01]]))
    else:
        kern = G.make_kern(["matern52", "white"], D)
        kern.setParams([np.sqrt(D), 1.0, 0.01])
    return kern, X, y


@pytest.mark.parametrize("N,D,kind", [(8192, 8, "rbf"), (16384, 16, "rbfard")])
def test_full_size_properties(N, D, kind):
    kern, X, y = _config(N, D, kind)
    gp = G.CGp(kern, X, y)
    g, ll = gp.logLikelihoodGradient()
    assert np.isfinite(ll) and np.isfinite(g).all()
    rng = np.random.default_rng(1)
    idx = rng.choice(N, 256, replace=False)
    K = gp.ctx.download(0)
    # (1) sampled K entries against the oracle's element formula
    ref = O.kern_compute(O.kern_from_trans([kind, "white"], kern.getTransParams(), D), X[np.sort(idx)])
    sub = K[np.ix_(np.sort(idx), np.sort(idx))]
    assert np.abs(sub - ref).max() < 1e-12
    # (2) factor / inverse residuals through random probes: K (Kinv v) = v, L L' v = K v
    L = gp.ctx.download(1)
    V = rng.standard_normal((N, 4))
    assert np.abs(L @ (L.T @ V) - K @ V).max() < 1e-9 * np.abs(K @ V).max()
    del L
    Kinv = gp.ctx.download(2)
    assert np.abs(K @ (Kinv @ V) - V).max() < 1e-7
    # (3) alpha and the quadratic form: K alpha = m
    alpha = gp.ctx.download(3)
    assert np.abs(K @ alpha - y).max() < 1e-8
    assert rel_err(float(np.sum(alpha * y)), gp._out[1]) < 1e-10
    # (4) logdet against numpy's slogdet of K is O(N^3) on the host: use the factor's diagonal instead
    # (5) white-noise gradient identity: dL/dsigma2 = -1/2 tr(Kinv) + 1/2 |alpha|^2, times gradfact = sigma2
    gw = (-0.5 * np.trace(Kinv) + 0.5 * float(np.sum(alpha * alpha))) * kern.getParam(kern.getNumParams() - 1)
    assert rel_err(g[-1], gw) < 1e-8
    # (6) directional finite difference of ll along a random direction in transformed-parameter space
    tp = kern.getTransParams()
    dvec = rng.standard_normal(tp.size)
    dvec /= np.linalg.norm(dvec)
    h = 1e-5
    gp.setOptParams(tp + h * dvec)
    lp = gp.logLikelihood()
    gp.setOptParams(tp - h * dvec)
    lm = gp.logLikelihood()
    assert abs((lp - lm) / (2 * h) - float(g @ dvec)) < 1e-4 * max(1.0, abs(float(g @ dvec)))


def test_wide_dynamic_range_evaluation_on_the_tensor_core_engine():
    """The tensor-core fp64 engine is exact relative to each operand ROW's largest entry (ozaki.cu), not element-wise.
    A kernel matrix whose rows span many decades -- poly + lin with large weights over inputs of very different norms,
    plus a tiny white term -- is the adversarial case for that; the engine is forced on for every product large enough
    (gpc_set_gemm_engine) and the full evaluation must still meet the 1e-8 bar against the oracle (N = 1536: the
    factorisation's GEMMs go through the engine)."""
    from gpc_b200._lib import lib
    rng = np.random.default_rng(21)
    N, D = 1536, 3
    X = rng.standard_normal((N, D)) * np.exp(rng.uniform(-3.0, 1.5, size=(N, 1)))   # row norms over 4.5 e-folds
    y = np.tanh(X[:, :1]) + 0.05 * rng.standard_normal((N, 1))
    types = ["poly", "lin", "rbf", "white"]
    tp = np.array([np.log(3.0), np.log(1e-4), np.log(5.0), np.log(4.0), 0.0, 0.0, 0.0])
    kern_o = O.kern_from_trans(types, tp, D)
    Kd = O.kern_compute(kern_o, X)
    # entries over 11 decades, a typical row over 5, cond(K) = 6e6 (so that 1e-8 is a fair bar for fp64)
    assert Kd.max() / np.abs(Kd).min() > 1e10 and np.median(np.abs(Kd).max(1) / np.abs(Kd).min(1)) > 1e5
    lib().gpc_set_gemm_engine(1, 8, 128, 128)      # tensor-core engine for everything from 128 x 128 x 128 up
    try:
        gp = G.CGp(G.make_kern(types, D, tp), X, y, bias=y.mean(0))
        g, ll = gp.logLikelihoodGradient()
    finally:
        lib().gpc_set_gemm_engine(1, 8, 256, 512)  # defaults (ozaki.cu)
    r = O.gp_loglik_grad(kern_o, X, y, bias=y.mean(0))
    assert rel_err(ll, r["ll"]) < 1e-8
    assert rel_err(g, r["g"]) < 1e-8


def test_gplvm_scg_trajectory_config5():
    """BASELINE config 5 (shortened to 12 SCG iterations for test time): CGplvm.fromData on oilTrain, SCG driven by
    the device evaluations, objective trajectory against the reference's own log (printed with 6 digits)."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    Y = np.load(os.path.join(root, "tests/golden/oil_train.npz"))["Y"]
    ref = json.load(open(os.path.join(root, "tests/golden/gplvm_c5_trajectory.json")))["objective"]
    lvm = G.CGplvm.fromData(G.make_kern(["rbf", "bias", "white"], 2, [0.0, 0.0, -2.0, -2.0]), Y, 2)
    log = []
    lvm.optimise(40, log=log)
    assert len(log) == 40
    for a, b in zip(log, ref[:40]):
        assert abs(a - b) <= 2e-5 * max(1.0, abs(b)), (a, b)
