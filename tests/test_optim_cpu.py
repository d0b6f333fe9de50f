"""The native SCG loop (gpc_scg_minimise, gpc_b200/csrc/host.cu) against the step-for-step Python restatement of
COptimisable::scgOptimise (COptimisable.cpp:246-396) on analytic objectives.  Host code only: runs without a GPU."""
import numpy as np
import pytest

from gpc_b200.optim import scgOptimise, scgOptimiseNative


class Model:
    """the COptimisable interface (COptimisable.h:15-239) around f(w), grad f(w)"""

    def __init__(self, f, g, w0):
        self.f, self.g, self.w = f, g, np.array(w0, dtype=np.float64)
        self.calls = 0

    def getOptParams(self):
        return self.w.copy()

    def setOptParams(self, w):
        self.w = np.array(w, dtype=np.float64)

    def computeObjectiveVal(self):
        return self.f(self.w)

    def computeObjectiveGradParams(self):
        self.calls += 1
        return self.g(self.w), self.f(self.w)


def rosen(w):
    return float(np.sum(100.0 * (w[1:] - w[:-1] ** 2) ** 2 + (1 - w[:-1]) ** 2))


def rosen_g(w):
    g = np.zeros_like(w)
    g[:-1] = -400.0 * w[:-1] * (w[1:] - w[:-1] ** 2) - 2 * (1 - w[:-1])
    g[1:] += 200.0 * (w[1:] - w[:-1] ** 2)
    return g


A = np.diag(np.arange(1.0, 7.0)) + 0.1 * np.ones((6, 6))
CASES = {
    "rosenbrock4": (rosen, rosen_g, [-1.2, 1.0, -0.5, 0.8], 150),
    "quadratic6": (lambda w: float(0.5 * w @ A @ w - w.sum()), lambda w: A @ w - 1.0, np.linspace(-1, 1, 6), 40),
    "one_parameter": (lambda w: float((w[0] - 3.0) ** 4 + w[0] ** 2), lambda w: np.array([4 * (w[0] - 3.0) ** 3 + 2 * w[0]]),
                      [0.0], 60),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_native_scg_reproduces_the_restatement(name):
    f, g, w0, iters = CASES[name]
    a, b = Model(f, g, w0), Model(f, g, w0)
    la, lb = [], []
    ita = scgOptimise(a, maxIters=iters, log=la)
    itb, evals = scgOptimiseNative(b, maxIters=iters, log=lb)
    assert ita == itb == len(lb)
    # the same steps; dot products are summed in a different order (numpy vs a plain loop, fused multiply-adds) and the
    # finite difference of step 2 (sigma = 1e-4/|p|) amplifies that rounding along the trajectory: the first dozen
    # iterations agree to ~1e-12, the end points to the accuracy the optimiser reaches
    k = min(12, len(la))
    assert np.allclose(la[:k], lb[:k], rtol=1e-9, atol=1e-12)
    assert np.allclose(la, lb, rtol=0.05, atol=1e-7)
    assert np.allclose(a.w, b.w, rtol=1e-3, atol=1e-6)
    assert lb[-1] <= lb[0]
    # one evaluation per distinct point: the restatement asks for the gradient again at every accepted point
    assert evals <= 2 * itb + 1 and b.calls == evals


def test_callback_failure_is_reported_not_thrown_across_the_abi():
    def bad(w):
        raise ValueError("objective failed")
    m = Model(bad, lambda w: w, [1.0, 2.0])
    with pytest.raises(ValueError):
        scgOptimiseNative(m, maxIters=5)
