"""The C++ host side of the drop-in on the B200 (north star: "host code stays C++ and calls through a thin C-ABI").

gpc_b200/cpp/CGpB200 : CGp and CGplvmB200 : CGplvm re-bind the reference's virtual entry points (logLikelihood,
logLikelihoodGradient, out) to libgpc_b200.so.  Two kinds of checks, both against the UNMODIFIED reference compiled by
oracle/build_ref.sh:
  * oracle/_ref/cgp_b200_check: reference class and drop-in class on the same data in ONE process -- log-likelihood,
    optimiser-space gradient (priors, transforms, learnt scales), predictions through out(), and the point the
    reference's own SCG optimiser reaches when it drives each class;
  * oracle/_ref/gp_l2: the reference's gp.cpp front-end compiled with `-include gp_dropin.h` (no source change) next to
    the plain OpenBLAS build: `gp learn` must print the same parameters and log-likelihood.
Tolerance: 1e-8 relative (BASELINE.json north_star), looser only where an optimiser trajectory amplifies rounding."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(os.path.dirname(HERE), "oracle", "_ref")
CHECK = os.path.join(REF, "cgp_b200_check")
TOL = 1e-8


def _check(*args):
    if not os.path.exists(CHECK):
        pytest.skip("oracle/_ref/cgp_b200_check not built (python __graft_entry__.py in the build container)")
    out = subprocess.run([CHECK] + [str(a) for a in args], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if not l.startswith("Warning:")]
    return json.loads("\n".join(lines))


def _rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), 1.0)))


def _assert_parity(r, opt_tol):
    assert r["on_device"] == 1
    assert r["evals_first"] == 1, "gradient + two likelihood calls at one point must cost ONE device evaluation"
    assert _rel(r["ll_ref"], r["ll_dev"]) <= TOL
    assert r["ll_dev"] == r["ll_dev_again"]
    assert _rel(r["g_ref"], r["g_dev"]) <= TOL, (r["g_ref"], r["g_dev"])
    assert _rel(r["out_ref"], r["out_dev"]) <= TOL
    assert _rel(r["std_ref"], r["std_dev"]) <= TOL
    if "out1_dev" in r:
        assert r["out1_dev"] == r["out_dev"]
    if "opt_ref" in r:
        assert r["opt_ll_ref"] > r["ll_ref"]
        assert _rel(r["opt_ref"], r["opt_dev"]) <= opt_tol, (r["opt_ref"], r["opt_dev"])
        assert _rel(r["opt_ll_ref"], r["opt_ll_dev"]) <= opt_tol


@pytest.mark.parametrize("N,D,d,kern,scale,prior,iters", [
    (40, 1, 1, "rbf,bias,white", 0, 0, 15),          # the shape of config 1
    (300, 3, 1, "rbf,lin,bias,white", 0, 0, 12),     # testGpftc's kernel
    (260, 4, 2, "rbfard,bias,white", 0, 0, 12),      # two outputs
    (220, 2, 1, "rbf,white", 1, 0, 12),              # learnt output scale (single output: the reference's limit)
    (200, 3, 1, "matern52,poly,white", 0, 1, 12),    # a gamma prior on the first parameter
    (500, 2, 3, "matern32,bias,white", 0, 1, 0),
])
def test_cgp_b200_matches_reference_cgp(N, D, d, kern, scale, prior, iters):
    _assert_parity(_check("gp", N, D, d, 7, kern, scale, prior, iters), opt_tol=1e-5)


@pytest.mark.parametrize("N,D,d,kern,approx,M,beta,iters", [
    (300, 2, 1, "rbf,bias,white", 1, 20, 50.0, 6),      # DTC (CGp::DTC = 1), then 6 SCG iterations driven by the reference's optimiser
    (400, 3, 2, "rbfard,white", 2, 30, 20.0, 0),        # FITC, two outputs, ARD
    (350, 2, 1, "rbf,lin,white", 4, 25, 30.0, 0),       # DTCVAR (no matern here: the reference's own dist2Row rounding gives
                                                        # NaN when an inducing input coincides with a data point, CKern.cpp:1839)
    (500, 2, 1, "rbf,white", 2, 140, 10.0, 0),          # FITC with M across a tile edge
])
def test_cgp_b200_sparse_matches_reference_cgp(N, D, d, kern, approx, M, beta, iters):
    """SURVEY 8(f) row 2 through the C++ host class: CGpB200 with a sparse approximation evaluates through gpc_sparse_eval /
    gpc_sparse_posterior; ll, the optimiser-space gradient [X_u][kernel][log beta], predictions and an SCG trajectory
    against the reference's CGp in the same process."""
    r = _check("sparsedev", N, D, d, 5, kern, approx, M, beta, iters)
    assert r["on_device"] == 1
    assert r["evals_first"] == 1
    assert _rel(r["ll_ref"], r["ll_dev"]) <= TOL
    assert r["ll_dev"] == r["ll_dev_again"]
    assert _rel(r["g_ref"], r["g_dev"]) <= 1e-7, (r["g_ref"], r["g_dev"])
    assert _rel(r["out_ref"], r["out_dev"]) <= TOL
    assert _rel(r["std_ref"], r["std_dev"]) <= TOL
    if "opt_ref" in r:
        assert r["opt_ll_ref"] > r["ll_ref"]
        assert _rel(r["opt_ll_ref"], r["opt_ll_dev"]) <= 1e-5


@pytest.mark.parametrize("N,q,d,kern,scale,prior,iters", [
    (120, 2, 5, "rbf,bias,white", 0, 0, 10),
    (200, 3, 4, "rbfard,white", 1, 0, 0),
    (150, 2, 6, "matern52,lin,white", 1, 1, 8),
])
def test_cgplvm_b200_matches_reference_cgplvm(N, q, d, kern, scale, prior, iters):
    _assert_parity(_check("gplvm", N, q, d, 11, kern, scale, prior, iters), opt_tol=1e-4)


def _write_svml(path, X, y):
    with open(path, "w") as f:
        for i in range(X.shape[0]):
            f.write("%.17g %s\n" % (y[i], " ".join("%d:%.17g" % (j + 1, X[i, j]) for j in range(X.shape[1]))))


def _learn(binary, data, model, iters, cwd):
    out = subprocess.run([binary, "-v", "2", "learn", "-#", str(iters), data, model], cwd=cwd, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    vals = {}
    for key in ("rbfinverseWidth", "rbfvariance", "biasvariance", "whitevariance", "Log likelihood"):
        m = re.findall(re.escape(key) + r":\s*([-+0-9.eE]+)", out.stdout)
        assert m, (key, out.stdout[-1500:])
        vals[key] = float(m[-1])
    return vals


@pytest.mark.parametrize("case", ["config1", "n700"])
def test_unmodified_gp_front_end_on_cgp_b200(tmp_path, case):
    """`gp learn` (gp.cpp:331-431) built on CGpB200 by the prefix header: the kernel matrix, its factorisation, inverse
    and gradients never exist on the host, the reference's SCG loop and model writer are untouched."""
    gp_cpu, gp_l2 = os.path.join(REF, "gp"), os.path.join(REF, "gp_l2")
    if not (os.path.exists(gp_cpu) and os.path.exists(gp_l2)):
        pytest.skip("oracle/_ref/gp and gp_l2 not built")
    if case == "config1":
        f = np.load(os.path.join(HERE, "golden", "gp_reference.npz"))
        X, y, iters = f["sinc_X"], np.asarray(f["sinc_y"]).ravel(), 100
    else:
        rng = np.random.default_rng(5)
        X = rng.standard_normal((700, 3))
        y = np.sin(X[:, 0]) * np.cos(0.5 * X[:, 1]) + 0.1 * rng.standard_normal(700)
        iters = 10
    data = str(tmp_path / "data.svml")
    _write_svml(data, X, y)
    a = _learn(gp_cpu, data, str(tmp_path / "m_cpu"), iters, str(tmp_path))
    b = _learn(gp_l2, data, str(tmp_path / "m_l2"), iters, str(tmp_path))
    for k in a:   # the CLI prints 6 significant digits
        assert abs(a[k] - b[k]) <= 2e-5 * max(1.0, abs(a[k])), (k, a[k], b[k])
    if case == "config1":   # README.md:103-106 of the reference
        assert abs(b["Log likelihood"] - 30.2364) < 1e-3
    # the model file written by the drop-in build is read back by the plain reference build (gp display)
    out = subprocess.run([gp_cpu, "display", str(tmp_path / "m_l2")], cwd=str(tmp_path), capture_output=True, text=True,
                         timeout=120)
    assert out.returncode == 0 and "rbfinverseWidth" in out.stdout


def test_unmodified_gplvm_front_end_on_cgplvm_b200_config5(tmp_path):
    """BASELINE config 5 through the reference's own front-end: `gplvm -v 3 -s 1 learn -# 40 oilTrain.svml` (gplvm.cpp
    compiled on CGplvmB200 by the prefix header; N=1000, q=2, rbf+bias+white, PCA initialisation and SCG are the
    reference's code) against the objective log of the unmodified OpenBLAS build (printed with 6 digits)."""
    exe = os.path.join(REF, "gplvm_l2")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/gplvm_l2 not built")
    Y = np.load(os.path.join(HERE, "golden", "oil_train.npz"))["Y"]
    ref = json.load(open(os.path.join(HERE, "golden", "gplvm_c5_trajectory.json")))["objective"]
    data = str(tmp_path / "oil.svml")
    with open(data, "w") as f:
        for i in range(Y.shape[0]):
            f.write("0 " + " ".join("%d:%.17g" % (j + 1, Y[i, j]) for j in range(Y.shape[1])) + "\n")
    out = subprocess.run([exe, "-v", "3", "-s", "1", "learn", "-#", "40", data, str(tmp_path / "oil.model")],
                         cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    errs = [float(v) for _, v in re.findall(r"Iteration:\s*(\d+)\s*Error:\s*([-+0-9.eE]+)", out.stdout)]
    assert len(errs) >= 40, out.stdout[-1500:]
    for a, b in zip(errs[:40], ref[:40]):
        assert abs(a - b) <= 2e-5 * max(1.0, abs(b)), (a, b)


def test_unmodified_gp_front_end_gnuplot_on_cgp_b200(tmp_path):
    """`gp gnuplot` (gp.cpp:567-906: the model file is read back into a default-constructed CGpB200, the prediction
    grid goes through out() -> gpc_kern_build, gpc_jitchol, gpc_posterior): plot data against the OpenBLAS build."""
    gp_cpu, gp_l2 = os.path.join(REF, "gp"), os.path.join(REF, "gp_l2")
    if not (os.path.exists(gp_cpu) and os.path.exists(gp_l2)):
        pytest.skip("oracle/_ref/gp and gp_l2 not built")
    f = np.load(os.path.join(HERE, "golden", "gp_reference.npz"))
    _write_svml(str(tmp_path / "sinc.svml"), f["sinc_X"], np.asarray(f["sinc_y"]).ravel())
    out = subprocess.run([gp_cpu, "-v", "1", "learn", "-#", "60", "sinc.svml", "model"], cwd=str(tmp_path),
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    for exe, tag in ((gp_cpu, "ref"), (gp_l2, "l2")):   # the SAME model file through both builds
        out = subprocess.run([exe, "gnuplot", "sinc.svml", "model", "plot_" + tag], cwd=str(tmp_path),
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    for part, tol in (("line_data", 2e-5), ("error_bar_data", 1e-8)):   # 6 and 18 printed digits
        a = np.loadtxt(str(tmp_path / ("plot_ref_%s.dat" % part)))
        b = np.loadtxt(str(tmp_path / ("plot_l2_%s.dat" % part)))
        assert a.shape == b.shape and a.size > 0
        assert _rel(a, b) <= tol, part


def test_download_accessors_of_cgp_b200():
    """CGpB200::downloadK / downloadInvK / downloadLcholK / downloadAlpha on the device-resident state: K against the
    reference's computeElement / diagComputeElement, K K^-1 = I, L L' = K with a zero strict upper triangle, alpha = K^-1 m."""
    if not os.path.exists(CHECK):
        pytest.skip("oracle/_ref/cgp_b200_check not built")
    out = subprocess.run([CHECK, "download", "300", "3", "2", "5", "rbf,lin,bias,white"], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stderr[-1500:]
    r = json.loads(out.stdout)
    assert r["refused_before_eval"] == 1
    assert r["rows"] == [300] * 4 and r["cols"] == [300, 300, 300, 2] and r["symmetric_flags"] == [1, 1, 0]
    assert r["err_K"] <= 1e-12 and r["err_KinvK"] <= 1e-8 and r["err_LLt"] <= 1e-10 and r["err_alpha"] <= 1e-8
    assert r["err_upper"] == 0.0


@pytest.mark.parametrize("N,D,N2,kern", [(300, 3, 170, "rbf,lin,bias,white"), (513, 2, 64, "rbfard,matern32,matern52,poly,white"),
                                         (100, 4, 200, "ratquad,rbf,white")])
def test_ccmpndkern_b200_compute_matches_reference(N, D, N2, kern):
    """The kernel-class seam of SURVEY 8(b): CKern::compute(K, X) and compute(K, X, X2) (CKern.h:128-157), called through
    the base class as CIvm.cpp:131 / CGp.cpp:540-545 / CGplvm.cpp:348 do.  A compound with a component outside the device
    path (ratquad) must fall through to the inherited loops, bit for bit."""
    r = _check("kern", N, D, N2, 3, kern)
    assert r["K_symmetric"] == 1
    if "ratquad" in kern:
        assert r["device_builds"] == 0 and r["K_maxdiff"] == 0.0 and r["K2_maxdiff"] == 0.0
    else:
        assert r["device_builds"] == 2
        assert r["K_maxdiff"] <= 1e-12 * max(1.0, r["K_max"])       # K entries to 1e-12
        assert r["K2_maxdiff"] <= 1e-12 * max(1.0, r["K_max"])
        assert r["clone_maxdiff"] <= 1e-12 * max(1.0, r["K_max"])


def _ivm_learn(exe, data, tmp_path):
    out = subprocess.run([exe, "-v", "1", "-s", "1", "learn", "-a", "100", "-k", "rbf", data, str(tmp_path / "m")],
                         cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    vals = {}
    for key in ("Active Set Size", "rbfinverseWidth", "rbfvariance", "biasvariance", "whitevariance", "Bias on process 0"):
        m = re.findall(re.escape(key) + r":\s*([-+0-9.eE]+)", out.stdout)
        assert m, (key, out.stdout[-1500:])
        vals[key] = float(m[-1])
    return vals


@pytest.mark.parametrize("exe", ["ivm_l2", "ivm_l1"])
def test_reference_ivm_front_end_on_the_device_classes(tmp_path, exe):
    """`ivm learn -a 100 -k rbf examples/unitsquaregp.svml` (README.md:234 of the reference) through (l2) the unmodified
    ivm.cpp compiled on CCmpndKernB200 -- its kernel matrices come from the device -- and (l1) the unmodified objects with
    the six CMatrix methods of gpc_b200/cpp/CMatrix_b200.cpp linked over the reference's weakened ones, against the plain
    OpenBLAS build.  The IVM's point selection is discrete: 6 printed digits with a small margin."""
    cpu, dev = os.path.join(REF, "ivm"), os.path.join(REF, exe)
    if not (os.path.exists(cpu) and os.path.exists(dev)):
        pytest.skip("oracle/_ref/%s not built" % exe)
    f = np.load(os.path.join(HERE, "golden", "unitsquaregp.npz"))
    data = str(tmp_path / "unitsquaregp.svml")
    _write_svml(data, f["X"], np.asarray(f["y"]).ravel())
    a, b = _ivm_learn(cpu, data, tmp_path), _ivm_learn(dev, data, tmp_path)
    assert a["Active Set Size"] == b["Active Set Size"] == 100
    for k in a:
        assert abs(a[k] - b[k]) <= 1e-4 * max(1.0, abs(a[k])), (k, a[k], b[k])


def test_reference_gp_front_end_on_level1_cmatrix(tmp_path):
    """INTEGRATION.md level 1 as compiled code: `gp learn` with CMatrix::potrf / potri / trsm / syrk / gemm / symv bound to
    gpc_d* (oracle/_ref/gp_l1) against the OpenBLAS build."""
    cpu, dev = os.path.join(REF, "gp"), os.path.join(REF, "gp_l1")
    if not (os.path.exists(cpu) and os.path.exists(dev)):
        pytest.skip("oracle/_ref/gp_l1 not built")
    rng = np.random.default_rng(21)
    X = rng.standard_normal((400, 2))
    y = np.sin(X[:, 0]) * np.cos(0.5 * X[:, 1]) + 0.1 * rng.standard_normal(400)
    data = str(tmp_path / "data.svml")
    _write_svml(data, X, y)
    res = []
    for exe in (cpu, dev):
        out = subprocess.run([exe, "-v", "2", "learn", "-#", "8", data, str(tmp_path / "m")], cwd=str(tmp_path),
                             capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        vals = {}
        for key in ("rbfinverseWidth", "rbfvariance", "biasvariance", "whitevariance", "Log likelihood"):
            m = re.findall(re.escape(key) + r":\s*([-+0-9.eE]+)", out.stdout)
            assert m, (key, out.stdout[-1500:])
            vals[key] = float(m[-1])
        res.append(vals)
    for k in res[0]:
        assert abs(res[0][k] - res[1][k]) <= 2e-5 * max(1.0, abs(res[0][k])), (k, res)
