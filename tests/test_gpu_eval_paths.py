"""GPU tests (-m gpu) of how gpc_eval schedules ONE evaluation (gpc_b200/csrc/api.cu): K built in the buffer that is
factored (the K buffer is rebuilt on demand), alpha = W'(W m) from the triangular inverse on the bulk stream, K^-1 left
lower-only on the device, the factorisation on a high-priority chain stream with its large side products on a
wave-limited bulk stream, the top-level look-ahead (TopFront) and the row-block pipeline (TopPipe).  Every variant
must give the numbers of the plain schedule and of the numpy oracle (CGp::logLikelihood / logLikelihoodGradient,
CGp.cpp:913-1144) to the 1e-8 bar of the north star; the scheduling switches are environment variables read when a
context is created / on first use, so the variants run in a subprocess each."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import gpc_b200 as G
from gpc_b200._lib import check, lib, ptr
from conftest import rel_err
from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem(N, D, d=1, seed=3):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, D))
    y = np.sin(X[:, :1]) @ np.ones((1, d)) + 0.1 * rng.standard_normal((N, d))
    kern = G.make_kern(["rbf", "white"], D)
    kern.setParams([1.0 / D, 1.0, 0.01])
    return kern, X, y - y.mean(axis=0)


def _eval(ctx, kern, want=0):
    arr, n, keep = kern._kcomps()
    out, g = np.zeros(3), np.zeros(kern.getNumParams())
    rc = check(lib().gpc_eval(ctx.handle, arr, n, want, ptr(out), ptr(g), None))
    assert rc == 0
    return out, g


@pytest.mark.parametrize("N,d", [(300, 1), (1100, 2), (2304, 1)])
def test_eval_leaves_consistent_state(N, d):
    """after gpc_eval: K (rebuilt on demand) is the kernel matrix, K^-1 downloads symmetric and inverts it, alpha from
    the triangular inverse equals the triangular solve (gpc_solve_alpha) and K alpha = m."""
    D = 5
    kern, X, y = _problem(N, D, d)
    ctx = G.DeviceContext(N, D, d)
    ctx.set_X(X)
    ctx.set_M(y)
    out, g = _eval(ctx, kern)
    Kref = O.kern_compute(O.kern_from_trans(["rbf", "white"], kern.getTransParams(), D), X)
    K = ctx.download(0)
    assert np.abs(K - Kref).max() < 1e-12
    Kinv = ctx.download(2)
    assert np.array_equal(Kinv, Kinv.T)
    assert np.abs(Kinv @ Kref - np.eye(N)).max() < 1e-8
    alpha = ctx.download(3)
    assert np.abs(Kref @ alpha - y).max() < 1e-9 * max(1.0, np.abs(y).max())
    quad = C.c_double(0)
    check(lib().gpc_solve_alpha(ctx.handle, C.byref(quad)))
    alpha2 = ctx.download(3)
    assert rel_err(alpha, alpha2) < 1e-9
    assert abs(quad.value - out[1]) < 1e-9 * max(1.0, abs(out[1]))
    # alpha through K^-1 (the reference's own route, CGp::updateAlpha) still works on the lower-only K^-1
    check(lib().gpc_alpha_from_inverse(ctx.handle, C.byref(quad)))
    assert rel_err(ctx.download(3), alpha) < 1e-8
    # a second evaluation (K lazily rebuilt in between) reproduces the first bit for bit
    out2, g2 = _eval(ctx, kern)
    assert np.array_equal(out, out2) and np.array_equal(g, g2)
    ctx.close()


_CHILD = r"""
import json, sys
import numpy as np
sys.path.insert(0, %(root)r)
import gpc_b200 as G
sys.path.insert(0, %(tests)r)
from test_gpu_eval_paths import _problem
out = {}
for N in %(sizes)r:
    kern, X, y = _problem(N, 6)
    gp = G.CGp(kern, X, y)
    g, ll = gp.logLikelihoodGradient()
    out[str(N)] = {"ll": float(ll), "g": [float(v) for v in g]}
print("RESULT " + json.dumps(out))
"""


def _run_variant(env_extra, sizes):
    env = dict(os.environ)
    env.update(env_extra)
    code = _CHILD % {"root": ROOT, "tests": os.path.join(ROOT, "tests"), "sizes": list(sizes)}
    p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


def test_scheduling_variants_agree_with_oracle():
    """plain schedule / chain priority + wave-limited bulk + look-ahead (default) / unlimited bulk without look-ahead /
    row-block pipeline on a narrower bulk partition: same numbers.
    Sizes cross the thresholds of the look-ahead (Np >= 2048) and of the pipeline (Np >= 1024)."""
    sizes = (900, 2100, 4200)
    variants = {
        "default": {},
        "plain": {"GPC_CHAIN_PRIO": "0"},
        "unlimited_bulk_no_front": {"GPC_BULK_SMS": "0", "GPC_TOP_FRONT": "0"},
        "pipeline_narrow": {"GPC_TOP_PIPE": "1", "GPC_BULK_SMS": "100"},
    }
    res = {k: _run_variant(v, sizes) for k, v in variants.items()}
    for N in sizes:
        kern, X, y = _problem(N, 6)
        r = O.gp_loglik_grad(O.kern_from_trans(["rbf", "white"], kern.getTransParams(), 6), X, y)
        for name, out in res.items():
            o = out[str(N)]
            assert rel_err(o["ll"], r["ll"]) < 1e-8, (name, N)
            assert rel_err(np.array(o["g"]), r["g"]) < 1e-8, (name, N)
            assert rel_err(o["ll"], res["plain"][str(N)]["ll"]) < 1e-10, (name, N)
            assert rel_err(np.array(o["g"]), np.array(res["plain"][str(N)]["g"])) < 1e-9, (name, N)
