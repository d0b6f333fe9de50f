"""GPU parity of SURVEY.md 8(f) row 2: the sparse approximations DTC / DTCVAR / FITC of CGp on the device
(gpc_sparse_* C ABI, gpc_b200/csrc/sparse.cu) and the cross-covariance gradient kernel under them.

Pins: the reference's MATLAB known answers matfiles/testGpdtc.mat / testGpfitc.mat (testGp.cpp:21-23, 98-150), outputs of
the compiled reference on seeded inputs (tests/golden/sparse_reference.npz: two outputs, ARD, poly; ll, the full
optimiser-space gradient [X_u][kernel][log beta] and the posterior), the kernel fixtures g4 / G2 / GD2 of
matfiles/*KernTest.mat (testKern.cpp:306-376), and the oracle (oracle/gp_sparse_oracle.py) on random problems."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpc_b200 as G  # noqa: E402
from conftest import SINGLE, rel_err  # noqa: E402
from gpc_b200.sparse import SparseGp  # noqa: E402
from oracle import gp_oracle as O  # noqa: E402
from oracle import gp_sparse_oracle as S  # noqa: E402

pytestmark = pytest.mark.gpu
MATCHTOL = 1e-10  # testKern.cpp / testMatrix.cpp absolute tolerance
SPARSE_TAGS = ["dtc_rbf", "dtc_ard2", "dtc_poly", "fitc_rbf", "fitc_ard2", "fitc_poly", "dtcvar_rbf", "dtcvar_ard2"]


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(ROOT, "tests", "golden", "sparse_reference.npz"))


def _model(ref, tag, approx):
    X, y, Xu = ref[tag + "_X"], ref[tag + "_y"], ref[tag + "_Xu"]
    D, M = X.shape[1], Xu.shape[0]
    types = [str(t) for t in ref[tag + "_types"]]
    P = sum(O.nparams(t, D) for t in types)
    params = ref[tag + "_params"]
    kern = G.make_kern(types, D, params[M * D:M * D + P])
    scale = [float(ref[tag + "_scale"])] if tag + "_scale" in ref else None
    gp = SparseGp(kern, X, y, Xu, float(ref[tag + "_beta"]), approx, bias=np.ravel(ref[tag + "_bias"]), scale=scale)
    np.testing.assert_allclose(gp.getOptParams(), params, rtol=0, atol=1e-12)   # [X_u][kernel][log beta], CGp.cpp:330-385
    return gp


@pytest.mark.parametrize("tag", SPARSE_TAGS)
def test_sparse_against_the_compiled_reference(ref, tag):
    gp = _model(ref, tag, str(ref[tag + "_approx"]))
    g, ll = gp.logLikelihoodGradient()
    assert abs(ll - float(ref[tag + "_ll"])) <= 1e-9 * max(1.0, abs(float(ref[tag + "_ll"])))
    assert rel_err(g, ref[tag + "_grads"]) <= 1e-8
    if tag + "_Xs" in ref:   # posterior mean / variance (CGp::posteriorMeanVar, sparse branches)
        mu, var = gp.posteriorMeanVar(ref[tag + "_Xs"])
        assert rel_err(mu, ref[tag + "_mu"]) <= 1e-8
        assert rel_err(var, ref[tag + "_var"]) <= 1e-8
    gp.close()


@pytest.mark.parametrize("tag,approx,gtol", [("testGpdtc", "dtc", 1e-5), ("testGpfitc", "fitc", 1e-6)])
def test_sparse_against_the_matlab_fixtures(ref, tag, approx, gtol):
    """testGp.cpp:22 fixtures.  Both are badly conditioned by construction (beta = 1000 with 50 inducing inputs,
    cond(A) = 1e9): the gradient is compared in norm like tests/test_oracle_sparse_cpu.py does for the oracle; the MATLAB
    `ll` lacks 1/2 d N log 2pi relative to the C++ value (same offset as testGpftc.mat)."""
    gp = _model(ref, tag, approx)
    g, ll = gp.logLikelihoodGradient()
    N, d = ref[tag + "_y"].shape
    ll_m = float(ref[tag + "_ll"])
    assert abs(ll + 0.5 * N * d * np.log(2 * np.pi) - ll_m) <= 1e-7 * abs(ll_m)
    gm = ref[tag + "_grads"]
    assert np.linalg.norm(g - gm) <= gtol * np.linalg.norm(gm)
    gp.close()


@pytest.mark.parametrize("approx", ["dtc", "fitc", "dtcvar"])
def test_sparse_random_vs_oracle_across_tile_edges(approx):
    """M and N that are not multiples of the 128 tile, several outputs, a compound kernel with every gradient rule"""
    rng = np.random.default_rng(17)
    N, M, D, d = 700, 150, 3, 3
    X = rng.standard_normal((N, D))
    Xu = X[rng.choice(N, M, replace=False)] + 0.05 * rng.standard_normal((M, D))
    y = np.sin(X[:, :1]) @ np.ones((1, d)) + 0.1 * rng.standard_normal((N, d))
    types = ["rbf", "matern32", "lin", "bias", "white"]
    tp = np.array([-0.3, 0.2, 0.4, -0.5, -1.5, -2.0, -3.0])
    beta = 25.0
    gp = SparseGp(G.make_kern(types, D, tp), X, y, Xu, beta, approx, bias=y.mean(0))
    g, ll = gp.logLikelihoodGradient()
    r = S.sparse_loglik_grad(O.kern_from_trans(types, tp, D), X, y, Xu, beta, approx, bias=y.mean(0))
    assert abs(ll - r["ll"]) <= 1e-9 * max(1.0, abs(r["ll"]))
    assert rel_err(g, r["g"]) <= 1e-7
    Xs = rng.standard_normal((37, D))
    mu, var = gp.posteriorMeanVar(Xs)
    mu_r, var_r = S.sparse_posterior(O.kern_from_trans(types, tp, D), X, y, Xu, beta, Xs, approx, bias=y.mean(0))
    assert rel_err(mu, mu_r) <= 1e-8 and rel_err(var, var_r) <= 1e-8
    gp.close()


@pytest.mark.parametrize("approx,N,M", [("dtc", 100000, 1024), ("fitc", 8000, 512)])
def test_sparse_large_n_vs_oracle(approx, N, M):
    """N = 100 000, M = 1024 -- the regime the approximation exists for (every M x M x N product on the tensor-core engine).
    The oracle still runs there in ~10 s of host time (it is O(N M^2)).  cond(A) is 1e7 at this size, so the inducing-input
    gradient, the most sensitive output, is held to 1e-6; ll and the kernel / beta gradients to the usual bars."""
    # (the oracle's FITC diagonal term materialises an N x N matrix: FITC is checked at N = 8000)
    rng = np.random.default_rng(3)
    D = 4
    X = rng.standard_normal((N, D))
    y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((N, 1))
    Xu = X[rng.choice(N, M, replace=False)].copy()
    types, tp, beta = ["rbf", "white"], np.array([np.log(0.5), 0.0, np.log(0.1)]), 10.0
    gp = SparseGp(G.make_kern(types, D, tp), X, y, Xu, beta, approx, bias=y.mean(0))
    g, ll = gp.logLikelihoodGradient()
    r = S.sparse_loglik_grad(O.kern_from_trans(types, tp, D), X, y, Xu, beta, approx, bias=y.mean(0))
    assert abs(ll - r["ll"]) <= 1e-9 * abs(r["ll"])
    assert rel_err(g[M * D:], r["g"][M * D:]) <= 1e-7
    assert rel_err(g[:M * D], r["g"][:M * D]) <= 1e-6
    gp.close()


# ---- the cross-covariance gradient kernel against the reference's MATLAB kernel fixtures -------------------------------
@pytest.mark.parametrize("name", SINGLE)
def test_cross_gradient_matlab_fixtures(kern_mat, name):
    """testKern.cpp:306-325 (g4: getGradTransParams(g, X, X2, covGrad2)), :327-362 (G2: getGradX) and :364-376 (GD2:
    getDiagGradX)."""
    f = kern_mat
    X, X2 = f[name + "_X"], f[name + "_X2"]
    D = X.shape[1]
    kern = G.make_kern([name], D, np.ravel(f[name + "_params"]))
    cg2 = f[name + "_covGrad2"]
    cg2 = cg2.toarray() if hasattr(cg2, "toarray") else np.asarray(cg2, dtype=np.float64)
    g4 = kern.getGradTransParams(X, cg2, X2=X2)
    assert np.abs(g4 - np.ravel(f[name + "_g4"])).max() < 1e-9   # sums of 20 000 terms: a decade above MATCHTOL
    # G2[i, j, :] = d k(X_i, X2_j) / d X_i for the first rows of X: contract with random weights, and pick single entries
    G2 = f[name + "_G2"]
    nr = G2.shape[0]
    rng = np.random.default_rng(4)
    W = np.zeros((X.shape[0], X2.shape[0]))
    W[:nr] = rng.standard_normal((nr, X2.shape[0]))
    _, gX = kern.getGradParams2(X, X2, W, want_gX=True)
    assert np.abs(gX[:nr] - np.einsum("ijk,ij->ik", G2, W[:nr])).max() < 1e-10 * X2.shape[0]
    assert np.abs(gX[nr:]).max() == 0.0
    for (i, j) in [(0, 0), (2, 117), (nr - 1, X2.shape[0] - 1)]:
        E = np.zeros_like(W)
        E[i, j] = 1.0
        _, gX = kern.getGradParams2(X, X2, E, want_gX=True)
        assert np.abs(gX[i] - G2[i, j]).max() < MATCHTOL
    # GD2: d k(x_i, x_i) / d x_i through the symmetric pass with covGrad = I
    _, gXd = kern.getGradParams(X, np.eye(X.shape[0]), want_gX=True)
    assert np.abs(gXd - np.asarray(f[name + "_GD2"], dtype=np.float64)).max() < MATCHTOL


def test_cross_gradient_compound_vs_oracle():
    rng = np.random.default_rng(9)
    X, X2 = rng.standard_normal((130, 3)), rng.standard_normal((257, 3))
    types = ["rbf", "rbfard", "matern32", "matern52", "lin", "poly", "bias", "white"]
    tp = rng.uniform(-1.0, 0.5, size=sum(O.nparams(t, 3) for t in types))
    W = rng.standard_normal((130, 257))
    kern = G.make_kern(types, 3, tp)
    g, gX = kern.getGradParams2(X, X2, W, want_gX=True)
    ok = O.kern_from_trans(types, tp, 3)
    assert rel_err(g, O.kern_grad_params(ok, X, W, X2=X2)) < 1e-10
    assert rel_err(gX, np.einsum("ijk,ij->ik", O.kern_gradX(ok, X, X2), W)) < 1e-10
