"""Sparse approximations of CGp -- DTC, DTCVAR, FITC with inducing inputs X_u -- host-side mirror of the gpc_sparse_* C
ABI (gpc_b200/csrc/sparse.cu; reference CGp.cpp:713-735, 766-861, 939-988, 1146-1413, 490-521, 584-599).  Everything
numerical runs in the library; this class keeps the reference's bookkeeping: m = (y - bias) / scale (CGp::updateM,
CGp.cpp:248-260), the optimiser's parameter order [X_u column-major][kernel transformed parameters][log beta]
(CGp.cpp:330-385) and the transforms.  No CPU path."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, fmat, lib, ptr

APPROX = {"dtc": 1, "fitc": 2, "dtcvar": 4}   # CGp::DTC, FITC, DTCVAR (CGp.h:13-19)


class SparseGp:
    def __init__(self, kern, X, y, Xu, beta, approx="dtc", bias=None, scale=None, device=0):
        X = fmat(X)
        y = fmat(np.asarray(y, dtype=np.float64).reshape(X.shape[0], -1))
        self.pkern = kern
        self.approx = approx.lower()
        self.N, self.D = X.shape
        self.d = y.shape[1]
        self.X_u = fmat(Xu).copy(order="F")
        self.M = self.X_u.shape[0]
        self.beta = float(beta)
        self.bias = np.zeros(self.d) if bias is None else np.asarray(bias, dtype=np.float64).reshape(self.d)
        self.scale = np.ones(self.d) if scale is None else np.asarray(scale, dtype=np.float64).reshape(self.d)
        self.m = fmat((y - self.bias[None, :]) / self.scale[None, :])
        self._h = C.c_void_p()
        check(lib().gpc_sparse_create(C.byref(self._h), device, APPROX[self.approx], self.N, self.M, self.D, self.d))
        check(lib().gpc_sparse_set_data(self._h, ptr(X), self.N, ptr(self.m), self.N))
        self._out = np.zeros(6)

    def close(self):
        if self._h:
            lib().gpc_sparse_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- optimiser interface (CGp::getOptParams / setOptParams, CGp.cpp:330-443) ------------------------------------
    def getOptNumParams(self):
        return self.M * self.D + self.pkern.getNumParams() + 1

    def getOptParams(self):
        return np.concatenate([self.X_u.reshape(-1, order="F"), self.pkern.getTransParams(), [np.log(self.beta)]])

    def setOptParams(self, p):
        p = np.asarray(p, dtype=np.float64)
        nx, nk = self.M * self.D, self.pkern.getNumParams()
        self.X_u = np.asfortranarray(p[:nx].reshape((self.M, self.D), order="F"))
        self.pkern.setTransParams(p[nx:nx + nk])
        self.beta = float(_lib.lib().gpc_transform_atox(1, float(p[nx + nk])))

    def logLikelihoodGradient(self):
        """(g, ll): g in the optimiser's order [X_u column-major][kernel transformed][log beta] (CGp.cpp:1016-1079)"""
        arr, n, keep = self.pkern._kcomps()
        gk = np.zeros(self.pkern.getNumParams())
        gXu = np.zeros((self.M, self.D), order="F")
        gb = C.c_double(0.0)
        rc = check(lib().gpc_sparse_eval(self._h, arr, n, ptr(self.X_u), self.M, self.beta, ptr(self._out), ptr(gk), ptr(gXu),
                                         C.byref(gb)))
        if rc > 0:
            raise _lib.MatrixNonPosDef(rc)
        g = np.concatenate([gXu.reshape(-1, order="F"), gk * self.pkern._gradfacts(), [gb.value * self.beta]])
        return g, float(self._out[0])

    def logLikelihood(self):
        return self.logLikelihoodGradient()[1]

    def posteriorMeanVar(self, Xs):
        """CGp::posteriorMeanVar, sparse branches (CGp.cpp:490-521, 584-599, 561-573, 618-626); call after an evaluation"""
        Xs = fmat(Xs)
        Ns = Xs.shape[0]
        arr, n, keep = self.pkern._kcomps()
        mu = np.zeros((Ns, self.d), order="F")
        var = np.zeros(Ns)
        check(lib().gpc_sparse_posterior(self._h, arr, n, ptr(Xs), Ns, Ns, ptr(mu), ptr(var)))
        return mu * self.scale[None, :] + self.bias[None, :], var[:, None] * (self.scale * self.scale)[None, :]
