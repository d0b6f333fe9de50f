"""Seam (1) of SURVEY.md 8(b) end to end: the UNMODIFIED reference `gp` front-end (compiled from /root/reference by
oracle/build_ref.sh) with its dpotrf_/dpotri_/dtrsm_/dsyrk_/dgemm_ calls resolved by gpc_b200/libgpc_lapack_shim.so
(-> libgpc_b200.so on the GPU), against the same binary on OpenBLAS: `gp learn` must arrive at the same kernel
parameters and log-likelihood."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(os.path.dirname(HERE), "oracle", "_ref")


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


def _both(tmp_path, X, y, iters):
    gp_cpu, gp_gpu = os.path.join(REF, "gp"), os.path.join(REF, "gp_b200")
    if not (os.path.exists(gp_cpu) and os.path.exists(gp_gpu)):
        pytest.skip("oracle/_ref/gp and gp_b200 not built (python __graft_entry__.py in the build container)")
    data = str(tmp_path / "data.svml")
    _write_svml(data, X, y)
    a = _learn(gp_cpu, data, str(tmp_path / "m_cpu"), iters, str(tmp_path))
    b = _learn(gp_gpu, data, str(tmp_path / "m_gpu"), iters, str(tmp_path))
    return a, b


def test_reference_gp_learn_on_the_b200_matches_openblas_config1(tmp_path):
    f = np.load(os.path.join(HERE, "golden", "gp_reference.npz"))
    X, y = f["sinc_X"], np.asarray(f["sinc_y"]).ravel()
    a, b = _both(tmp_path, X, y, 20)
    for k in a:   # the CLI prints 6 significant digits
        assert abs(a[k] - b[k]) <= 2e-5 * max(1.0, abs(a[k])), (k, a[k], b[k])


def test_reference_gp_learn_on_the_b200_matches_openblas_n600(tmp_path):
    rng = np.random.default_rng(21)
    X = rng.standard_normal((600, 2))
    y = np.sin(X[:, 0]) * np.cos(0.5 * X[:, 1]) + 0.1 * rng.standard_normal(600)
    a, b = _both(tmp_path, X, y, 8)
    for k in a:
        assert abs(a[k] - b[k]) <= 2e-5 * max(1.0, abs(a[k])), (k, a[k], b[k])
