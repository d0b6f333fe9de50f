"""Pins the numpy oracle (oracle/gp_oracle.py) against (1) the reference's MATLAB golden vectors and
(2) outputs of the unmodified reference compiled here, both committed under tests/golden/.
Tolerances follow the reference's own MATCHTOL = 1e-10 absolute (ndlutil.h:33) unless stated."""
import numpy as np
import pytest

from conftest import CASES, SINGLE, rel_err
from oracle import gp_oracle as O

MATCHTOL = 1e-10


@pytest.mark.parametrize("name", SINGLE)
def test_kernel_matlab_fixtures(kern_mat, name):
    """testKern.cpp:236-376 restated: compute, cross compute, diag, gradients, gradX."""
    f = kern_mat
    X, X2, tp = f[name + "_X"], f[name + "_X2"], f[name + "_params"]
    kern = O.kern_from_trans([name], tp, X.shape[1])
    assert np.abs(O.kern_compute(kern, X) - f[name + "_K2"]).max() < MATCHTOL
    assert np.abs(O.kern_cross(kern, X, X2) - f[name + "_K4"]).max() < MATCHTOL
    assert np.abs(O.kern_diag(kern, X) - f[name + "_k2"]).max() < MATCHTOL
    assert np.abs(O.kern_grad_trans_params(kern, X, f[name + "_covGrad"]) - f[name + "_g2"]).max() < MATCHTOL
    assert np.abs(O.kern_grad_trans_params(kern, X, f[name + "_covGrad2"], X2) - f[name + "_g4"]).max() < MATCHTOL
    assert np.abs(O.kern_gradX(kern, X[:6], X2) - f[name + "_G2"]).max() < MATCHTOL
    assert np.abs(O.kern_diagGradX(kern, X) - f[name + "_GD2"]).max() < MATCHTOL or name in ("lin", "poly")


@pytest.mark.parametrize("tag", list(CASES))
def test_compound_vs_reference(rand_ref, tag):
    f, types = rand_ref, CASES[tag]
    X, X2, tp = f[tag + "_X"], f[tag + "_X2"], f[tag + "_tparams"]
    kern = O.kern_from_trans(types, tp, X.shape[1])
    assert np.abs(np.concatenate([p for _, p in kern]) - f[tag + "_params"]).max() < 1e-14
    assert np.abs(O.kern_compute(kern, X) - f[tag + "_K"]).max() < 1e-11
    assert np.abs(O.kern_cross(kern, X, X2) - f[tag + "_Kx"]).max() < 1e-11
    assert np.abs(O.kern_diag(kern, X2) - f[tag + "_kdiag"]).max() < 1e-11
    assert rel_err(O.kern_grad_trans_params(kern, X, f[tag + "_covGrad"]), f[tag + "_g"]) < 1e-10
    assert rel_err(O.kern_grad_trans_params(kern, X, f[tag + "_covGrad2"], X2), f[tag + "_g2"]) < 1e-10
    assert np.abs(O.kern_gradX(kern, X[:5], X2) - f[tag + "_gradX"]).max() < 1e-10
    assert np.abs(O.kern_diagGradX(kern, X) - f[tag + "_diagGradX"]).max() < 1e-10


@pytest.mark.parametrize("tag", list(CASES))
def test_gp_eval_vs_reference(rand_ref, tag):
    """ll, gradient, posterior: 1e-8 relative (north_star tolerance)."""
    f, types = rand_ref, CASES[tag]
    X, y = f[tag + "_X"], f[tag + "_y"]
    kern = O.kern_from_trans(types, f[tag + "_tparams"], X.shape[1])
    r = O.gp_loglik_grad(kern, X, y, bias=f[tag + "_bias"], scale=f[tag + "_scale"])
    assert rel_err(r["ll"], f[tag + "_ll"]) < 1e-8
    assert rel_err(r["g"], f[tag + "_gll"]) < 1e-8
    mu, var = O.gp_posterior(kern, X, y, f[tag + "_X2"], bias=f[tag + "_bias"], scale=f[tag + "_scale"])
    assert rel_err(mu, f[tag + "_mu"]) < 1e-8
    assert rel_err(var, f[tag + "_var"]) < 1e-8


def test_gp_ftc_matlab(gp_ref):
    """matfiles/testGpftc.mat (testGp.cpp:118-150; 'ftc' is commented out there because CGp.cpp:1012 subtracts
    d N/2 log 2pi that the MATLAB ll omits): compare ll + d N/2 log 2pi."""
    f = gp_ref
    kern = O.kern_from_trans(["rbf", "lin", "bias", "white"], f["ftc_params"], 2)
    r = O.gp_loglik_grad(kern, f["ftc_X"], f["ftc_y"], bias=f["ftc_bias"])
    N = f["ftc_X"].shape[0]
    assert abs(r["ll"] + N * O.HALFLOGTWOPI - float(f["ftc_ll_matlab"])) < 1e-9
    assert np.abs(r["g"] - f["ftc_grads_matlab"]).max() < 1e-9
    assert rel_err(r["ll"], float(f["ftc_ll_ref"])) < 1e-12
    mu, var = O.gp_posterior(kern, f["ftc_X"], f["ftc_y"], f["ftc_Xs"], bias=f["ftc_bias"])
    assert rel_err(mu, f["ftc_mu_ref"]) < 1e-9 and rel_err(var, f["ftc_var_ref"]) < 1e-9


def test_sinc_config1(gp_ref):
    """BASELINE config 1: examples/sinc.svml, rbf+bias+white at gp.cpp defaults; SURVEY 8(c) known answers."""
    f = gp_ref
    kern = O.kern_from_trans(["rbf", "bias", "white"], f["sinc_params"], 1)
    r = O.gp_loglik_grad(kern, f["sinc_X"], f["sinc_y"], bias=f["sinc_bias"])
    assert abs(r["ll"] - (-28.2080301154265)) < 1e-9
    assert np.abs(r["g"] - np.array([-4.6975707918852, -10.9705430321513, -0.33705464079873, -7.92625419624426])).max() < 1e-9
    assert rel_err(r["g"], f["sinc_g_ref"]) < 1e-10
    mu, var = O.gp_posterior(kern, f["sinc_X"], f["sinc_y"], f["sinc_Xs"], bias=f["sinc_bias"])
    assert abs(mu[0, 0] - 0.940483662893009) < 1e-9 and abs(var[0, 0] - 0.192171243413303) < 1e-9
    assert rel_err(mu, f["sinc_mu_ref"]) < 1e-9 and rel_err(var, f["sinc_var_ref"]) < 1e-9


def test_dense_primitives_matlab(matrix_mat):
    """testMatrix.cpp testCholesky :206-236, testInv :187-205."""
    f = matrix_mat
    L, info = O.chol_lower(f["choleskyMatrixTest_C"])
    assert info == 0
    assert np.abs(L - f["choleskyMatrixTest_L"]).max() < MATCHTOL
    assert np.abs(L.T - f["choleskyMatrixTest_U"]).max() < MATCHTOL
    A = f["invMatrixTest_A"]
    La, info = O.chol_lower(A)
    if info == 0:
        assert np.abs(O.pdinv(La) - f["invMatrixTest_Ainv"]).max() < 1e-8
    B, Lt = f["trsmMatrixTest_B"], f["trsmMatrixTest_L"]
    Xs = O.solve_lower(Lt, B)
    assert np.abs(Lt @ Xs - B).max() < 1e-10


def test_jitchol_vs_reference(rand_ref):
    """CMatrix::jitChol CMatrix.cpp:767-804 on a rank-deficient matrix: same jitter, same mutated A."""
    f = rand_ref
    L, jit, Aout = O.jit_chol(f["jit_A"])
    assert rel_err(jit, float(f["jit_val"])) < 1e-12
    assert np.abs(Aout - f["jit_Aout"]).max() < 1e-10
    assert np.abs(L @ L.T - Aout).max() < 1e-8


def test_chol_info():
    A = np.eye(5)
    A[3, 3] = -1.0
    _, info = O.chol_lower(A)
    assert info == 4


def test_gplvm_vs_reference(rand_ref):
    f = rand_ref
    kern = O.kern_from_trans(["rbf", "bias", "white"], f["lvm_tparams"], 2)
    r = O.gplvm_loglik_grad(kern, f["lvm_X"], f["lvm_m"])
    assert rel_err(r["ll"], float(f["lvm_ll"])) < 1e-8
    assert rel_err(r["g"], f["lvm_g"]) < 1e-7
    kern = O.kern_from_trans(["rbf", "lin", "matern32", "white"], f["lvm2_tparams"], 2)
    r = O.gplvm_loglik_grad(kern, f["lvm2_X"], f["lvm_m"])
    assert rel_err(r["ll"], float(f["lvm2_ll"])) < 1e-8
    assert rel_err(r["g"], f["lvm2_g"]) < 1e-7
