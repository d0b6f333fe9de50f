"""Host-side mirror of CGp (FTC) and CGplvm (plain FTC) -- CGp.h / CGp.cpp, CGplvm.h / CGplvm.cpp.
Method names, parameter order and return conventions follow the reference; every O(N^2)/O(N^3) step runs in
libgpc_b200.so on the device-resident state of a gpc_ctx."""
import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import check, fmat, lib, ptr
from .kern import DeviceContext

HALFLOGTWOPI = 0.5 * math.log(2.0 * math.pi)  # ndlutil.h:39


class CGp:
    """Exact (FTC) Gaussian process: CGp(pkern, pnoise, pX, FTC, ...) with a Gaussian noise whose targets are y.

    logLikelihood()            CGp.cpp:913-1013
    logLikelihoodGradient(g)   CGp.cpp:1016-1079  (kernel transformed-parameter gradients, component order)
    posteriorMeanVar(Xs)       CGp.cpp:642-663
    """

    def __init__(self, kern, X, y, bias=None, scale=None, device=0):
        self.pkern = kern
        self.X = fmat(X)
        self.y = fmat(y)
        N, d = self.y.shape
        assert self.X.shape[0] == N, "CGp: X and y must have the same number of rows"
        self.bias = np.zeros(d) if bias is None else np.asarray(bias, dtype=np.float64).reshape(d)
        self.scale = np.ones(d) if scale is None else np.asarray(scale, dtype=np.float64).reshape(d)
        self.ctx = DeviceContext(N, self.X.shape[1], d, device)
        self.ctx.set_X(self.X)
        self.updateM()
        self.KupToDate = False
        self._out = np.zeros(3)
        self._g = None

    @classmethod
    def fromModelFile(cls, path, X, y, device=0):
        """readGpFromFile + the caller's data, as gp.cpp:486-490 / 562-622 do for relearn, display and gnuplot: the
        file's kernel, scale and bias (native reader, gpc_gp_model_read) around X and y."""
        from . import io
        m = io.read_gp_model(path)
        return cls(io.kern_from_model(m), X, y, bias=m["bias"], scale=m["scale"], device=device)

    def writeModelFile(self, path, comment=""):
        """writeGpToFile (CGp.cpp:1689-1692): a model file the reference's gp display / gnuplot / relearn read."""
        from . import io
        io.write_gp_model(path, io.model_from_gp(self), comment)

    # --- CGp::updateM (CGp.cpp:248-260)
    def updateM(self):
        self.m = fmat((self.y - self.bias[None, :]) / self.scale[None, :])
        self.ctx.set_M(self.m)
        self.KupToDate = False

    def setBias(self, b):
        self.bias = np.asarray(b, dtype=np.float64).reshape(-1)

    def setScale(self, s):
        self.scale = np.asarray(s, dtype=np.float64).reshape(-1)

    def getNumData(self):
        return self.X.shape[0]

    def getOutputDim(self):
        return self.y.shape[1]

    # --- optimiser interface (CGp.cpp:330-443): FTC, fixed X, no learnt scales => kernel trans-params only
    def getOptNumParams(self):
        return self.pkern.getNumParams()

    def getOptParams(self):
        return self.pkern.getTransParams()

    def setOptParams(self, p):
        self.pkern.setTransParams(p)
        self.KupToDate = False

    def _eval(self):
        """K build -> jitChol -> K^-1 -> alpha -> gradient: one fused device evaluation (gpc_eval)."""
        if self.KupToDate:
            return
        arr, n, keep = self.pkern._kcomps()
        g = np.zeros(self.pkern.getNumParams())
        rc = check(lib().gpc_eval(self.ctx.handle, arr, n, 0, ptr(self._out), ptr(g), None))
        if rc > 0:
            raise _lib.MatrixNonPosDef(rc)
        self._g = g * self.pkern._gradfacts()
        self.KupToDate = True

    def logLikelihood(self):
        self._eval()
        logdet, quad = self._out[0], self._out[1]
        N, d = self.getNumData(), self.getOutputDim()
        return -0.5 * (quad + d * logdet) - d * N * HALFLOGTWOPI

    def logLikelihoodGradient(self):
        """returns (g, ll) -- the reference fills g and returns ll."""
        self._eval()
        return self._g.copy(), self.logLikelihood()

    def computeObjectiveVal(self):
        return -self.logLikelihood()

    def computeObjectiveGradParams(self):
        g, ll = self.logLikelihoodGradient()
        return -g, -ll

    def optimise(self, iters=1000, verbosity=0, log=None):
        """CGp::optimise / CGplvm::optimise -> runDefaultOptimiser -> scgOptimise (CGp.cpp:1537-1553)."""
        from .optim import scgOptimise
        return scgOptimise(self, maxIters=iters, verbosity=verbosity, log=log)

    def optimiseNative(self, iters=1000, paramTol=1e-6, objectiveTol=1e-6, log=None):
        """The same SCG loop run natively inside the library (gpc_gp_optimise_scg): no Python between evaluations.
        Returns (iterations, device evaluations)."""
        arr, n, keep = self.pkern._kcomps()
        trace = np.zeros(max(int(iters), 1))
        it, ev = C.c_int(0), C.c_int(0)
        rc = check(lib().gpc_gp_optimise_scg(self.ctx.handle, arr, n, int(iters), paramTol, objectiveTol, ptr(trace),
                                             C.byref(it), C.byref(ev)))
        # the library wrote the optimum into the parameter arrays it was given: copy back into the kernel objects
        for k, pvals in zip(self.pkern._components(), keep):
            k.params = np.array(pvals, dtype=np.float64)
        self.KupToDate = False
        if log is not None:
            log.extend(trace[:it.value].tolist())
        if rc > 0:
            raise _lib.MatrixNonPosDef(rc)
        return it.value, ev.value

    def posteriorMeanVar(self, Xs):
        """mu, var at Xs with output scale/bias applied (CGp.cpp:561-573, 618-623)."""
        self._eval()
        Xs = fmat(Xs)
        Ns, d = Xs.shape[0], self.getOutputDim()
        # the reference's posterior alpha comes from the triangular factor (CGp::updateAlpha, CGp.cpp:469-484)
        quad = C.c_double(0)
        check(lib().gpc_solve_alpha(self.ctx.handle, C.byref(quad)))
        arr, n, keep = self.pkern._kcomps()
        mu = np.zeros((Ns, d), order="F")
        var = np.zeros((Ns, d), order="F")
        check(lib().gpc_posterior(self.ctx.handle, arr, n, ptr(Xs), Ns, Ns, ptr(mu), ptr(var)))
        mu = mu * self.scale[None, :] + self.bias[None, :]
        var = var * (self.scale * self.scale)[None, :]
        return mu, var

    def timings(self):
        return self.ctx.last_timings()


class CGplvm:
    """GP-LVM, plain FTC with the Gaussian latent prior (CGplvm.cpp:493-716).  Optimiser parameter order is
    [kernel trans-params][X column-major] (CGplvm.cpp:257-290).  m = (Y - bias)/scale as CScaleNoise::updateSites
    leaves it (CNoise.cpp:710-721)."""

    @classmethod
    def fromData(cls, kern, Y, latentDim=2, device=0, centreData=True, scaleData=False):
        """What `gplvm learn` builds (gplvm.cpp:504-522): CScaleNoise targets m = (Y - bias)/scale with bias = column
        means when centring (the CLI default) and scale = population std only with -S (CNoise.cpp:576-587, 710-721;
        gplvm.cpp:506-513), and the PCA initialisation of X (CGplvm::initXpca, CGplvm.cpp:157-222):
        X = m U_q diag(lambda_q)^-1/2 with (lambda, U) the leading eigenpairs of cov(m), then centred."""
        Y = np.asarray(Y, dtype=np.float64)
        scale = np.sqrt(Y.var(axis=0)) if scaleData else np.ones(Y.shape[1])
        scale[scale < np.finfo(np.float64).eps] = np.finfo(np.float64).eps
        bias = Y.mean(axis=0) if centreData else np.zeros(Y.shape[1])
        m = (Y - bias[None, :]) / scale[None, :]
        ymean = m.mean(axis=0)
        cov = m.T @ m / m.shape[0] - np.outer(ymean, ymean)
        ev, U = np.linalg.eigh(cov)
        Winv = np.stack([U[:, -1 - i] / np.sqrt(ev[-1 - i]) for i in range(latentDim)], axis=1)
        X0 = m @ Winv
        X0 = X0 - X0.mean(axis=0)[None, :]
        return cls(kern, m, X0, device)

    def __init__(self, kern, m, X0, device=0):
        self.pkern = kern
        self.m = fmat(m)
        self.X = fmat(X0).copy(order="F")
        N, d = self.m.shape
        self.ctx = DeviceContext(N, self.X.shape[1], d, device)
        self.KupToDate = False
        self._out = np.zeros(3)

    def getOptNumParams(self):
        return self.pkern.getNumParams() + self.X.size

    def getOptParams(self):
        return np.concatenate([self.pkern.getTransParams(), self.X.reshape(-1, order="F")])

    def setOptParams(self, p):
        nk = self.pkern.getNumParams()
        self.pkern.setTransParams(p[:nk])
        self.X = np.asfortranarray(np.asarray(p[nk:], dtype=np.float64).reshape(self.X.shape, order="F"))
        self.KupToDate = False

    def _eval(self):
        if self.KupToDate:
            return
        self.ctx.set_X(self.X)
        self.ctx.set_M(self.m)
        arr, n, keep = self.pkern._kcomps()
        g = np.zeros(self.pkern.getNumParams())
        gX = np.zeros(self.X.shape, order="F")
        rc = check(lib().gpc_eval(self.ctx.handle, arr, n, 1, ptr(self._out), ptr(g), ptr(gX)))
        if rc > 0:
            raise _lib.MatrixNonPosDef(rc)
        self._gk = g * self.pkern._gradfacts()
        self._gX = gX - self.X  # latent prior gradient -X (CGplvm.cpp:672-681)
        self.KupToDate = True

    def logLikelihood(self):
        self._eval()
        d = self.m.shape[1]
        # no -dN/2 log 2pi here, unlike CGp (CGplvm.cpp:547-552)
        return -0.5 * (self._out[1] + d * self._out[0] + float(np.sum(self.X * self.X)))

    def logLikelihoodGradient(self):
        self._eval()
        return np.concatenate([self._gk, self._gX.reshape(-1, order="F")]), self.logLikelihood()

    def computeObjectiveVal(self):
        return -self.logLikelihood()

    def computeObjectiveGradParams(self):
        g, ll = self.logLikelihoodGradient()
        return -g, -ll

    def optimise(self, iters=1000, verbosity=0, log=None):
        """CGp::optimise / CGplvm::optimise -> runDefaultOptimiser -> scgOptimise (CGp.cpp:1537-1553)."""
        from .optim import scgOptimise
        return scgOptimise(self, maxIters=iters, verbosity=verbosity, log=log)
