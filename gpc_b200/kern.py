"""Host-side mirror of the reference's kernel classes (CKern.h / CKern.cpp) for the in-scope components.
Same class names, parameter order, transforms and method names as the reference so that the parity tests read
like testKern.cpp; all O(N^2) work is done by libgpc_b200.so (no numpy fallback)."""
import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import KComp, check, fmat, lib, ptr

WHITE, BIAS, RBF, RBFARD, MATERN32, MATERN52, LIN, POLY = range(8)
TYPE_NAMES = {WHITE: "white", BIAS: "bias", RBF: "rbf", RBFARD: "rbfard", MATERN32: "matern32",
              MATERN52: "matern52", LIN: "lin", POLY: "poly"}


class DeviceContext:
    """Owns a gpc_ctx (device-resident X, m, K, L, K^-1, alpha of one model)."""

    def __init__(self, Nmax, Dmax, dout_max=1, device=0):
        self._h = C.c_void_p()
        check(lib().gpc_ctx_create(C.byref(self._h), int(device), int(Nmax), int(Dmax), int(dout_max)))
        self.Nmax, self.Dmax, self.dmax, self.device = Nmax, Dmax, dout_max, device
        self.N = self.D = self.d = 0

    def close(self):
        if getattr(self, "_h", None):
            lib().gpc_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        if not self._h:
            raise _lib.GpcError("context closed")
        return self._h

    def set_X(self, X):
        X = fmat(X)
        self._X = X
        check(lib().gpc_set_X(self.handle, ptr(X), X.shape[0], X.shape[1], X.shape[0]))
        self.N, self.D = X.shape

    def set_M(self, M):
        M = fmat(M)
        self._M = M
        check(lib().gpc_set_M(self.handle, ptr(M), M.shape[0], M.shape[1], M.shape[0]))
        self.d = M.shape[1]

    def set_X_ptr(self, addr, N, D, ld):
        check(lib().gpc_set_X(self.handle, C.c_void_p(addr), N, D, ld))
        self.N, self.D = N, D

    def set_M_ptr(self, addr, N, d, ld):
        check(lib().gpc_set_M(self.handle, C.c_void_p(addr), N, d, ld))
        self.d = d

    def sync(self):
        check(lib().gpc_ctx_sync(self.handle))

    def stream(self):
        return lib().gpc_ctx_get_stream(self.handle)

    def set_stream(self, s):
        check(lib().gpc_ctx_set_stream(self.handle, C.c_void_p(s)))

    def launch_count(self):
        return int(lib().gpc_ctx_launch_count(self.handle))

    def download(self, which):
        cols = self.N if which < 3 else self.d
        out = np.zeros((self.N, cols), order="F")
        check(lib().gpc_download(self.handle, which, ptr(out), self.N))
        return out

    def last_timings(self):
        t = np.zeros(6)
        check(lib().gpc_last_timings(self.handle, ptr(t)))
        return dict(kbuild=t[0], potrf=t[1], inverse=t[2], alpha=t[3], grad=t[4], total=t[5])


class CKern:
    """Base of the mirrored kernel classes (CKern.h:36-355)."""
    type_code = None
    type_name = None

    def __init__(self, inDim_or_X):
        self.inputDim = int(inDim_or_X if np.isscalar(inDim_or_X) else np.asarray(inDim_or_X).shape[1])
        self.degree = 2.0
        self.setInitParam()

    # -- parameters -------------------------------------------------------------------------------------
    def getNumParams(self):
        return lib().gpc_kern_nparams(self.type_code, self.inputDim)

    def getInputDim(self):
        return self.inputDim

    def getType(self):
        return self.type_name

    def setParam(self, val, paramNo):
        self.params[paramNo] = float(val)

    def getParam(self, paramNo):
        return float(self.params[paramNo])

    def getParams(self):
        return self.params.copy()

    def setParams(self, p):
        p = np.asarray(p, dtype=np.float64).ravel()
        assert p.size == self.getNumParams()
        self.params = p.copy()

    def _transform(self, i):
        return lib().gpc_kern_transform(self.type_code, int(i))

    def getTransParam(self, i):
        return lib().gpc_transform_xtoa(self._transform(i), self.params[i])

    def setTransParam(self, val, i):
        self.params[i] = lib().gpc_transform_atox(self._transform(i), float(val))

    def getTransParams(self):
        return np.array([self.getTransParam(i) for i in range(self.getNumParams())])

    def setTransParams(self, tp):
        """CTransformable::setTransParams (CTransform.h:281-300)."""
        tp = np.asarray(tp, dtype=np.float64).ravel()
        assert tp.size == self.getNumParams(), "setTransParams(): Dimension match check failed"
        for i in range(tp.size):
            self.setTransParam(tp[i], i)

    def _gradfacts(self):
        return np.array([lib().gpc_transform_gradfact(self._transform(i), self.params[i])
                         for i in range(self.getNumParams())])

    # -- component list (compound kernels override) -------------------------------------------------------
    def _components(self):
        return [self]

    def _kcomps(self):
        comps = self._components()
        arr = (KComp * len(comps))()
        keep = []
        for i, k in enumerate(comps):
            p = np.ascontiguousarray(k.params, dtype=np.float64)
            keep.append(p)
            arr[i].type = k.type_code
            arr[i].nparams = p.size
            arr[i].params = p.ctypes.data_as(_lib.c_double_p)
            arr[i].degree = k.degree
        return arr, len(comps), keep

    # -- computations (device) ----------------------------------------------------------------------------
    def _ctx_for(self, X):
        X = fmat(X)
        ctx = DeviceContext(X.shape[0], X.shape[1], 1)
        ctx.set_X(X)
        return ctx

    def compute(self, X, X2=None):
        """CKern::compute(K, X) (CKern.h:128-144) or compute(K, X, X2) (:146-157); returns K."""
        ctx = self._ctx_for(X)
        try:
            arr, n, keep = self._kcomps()
            if X2 is None:
                check(lib().gpc_kern_build(ctx.handle, arr, n))
                return ctx.download(0)
            X2 = fmat(X2)
            K = np.zeros((ctx.N, X2.shape[0]), order="F")
            check(lib().gpc_kern_cross(ctx.handle, arr, n, ptr(X2), X2.shape[0], X2.shape[0], ptr(K), ctx.N))
            return K
        finally:
            ctx.close()

    def diagCompute(self, X):
        """CKern::diagCompute (CKern.h:50-56)."""
        ctx = self._ctx_for(X)
        try:
            arr, n, keep = self._kcomps()
            X = fmat(X)
            d = np.zeros(X.shape[0])
            check(lib().gpc_kern_diag(ctx.handle, arr, n, ptr(X), X.shape[0], X.shape[0], ptr(d)))
            return d
        finally:
            ctx.close()

    def getGradParams(self, X, covGrad, want_gX=False):
        """CKern::getGradParams(g, X, covGrad) (CKern.h:187-197 + overrides), natural parameters."""
        ctx = self._ctx_for(X)
        try:
            arr, n, keep = self._kcomps()
            cg = fmat(covGrad)
            g = np.zeros(self.getNumParams())
            gX = np.zeros((ctx.N, ctx.D), order="F") if want_gX else None
            check(lib().gpc_kern_grad(ctx.handle, arr, n, ptr(cg), cg.shape[0], ptr(g), ptr(gX) if want_gX else None))
            return (g, gX) if want_gX else g
        finally:
            ctx.close()

    def getGradTransParams(self, X, covGrad, X2=None):
        """CKern::getGradTransParams (CKern.cpp:50-63): gradient w.r.t. the transformed parameters; with X2 the
        cross-covariance form getGradTransParams(g, X, X2, covGrad2) (testKern.cpp:306-325)."""
        if X2 is not None:
            return self.getGradParams2(X, X2, covGrad) * self._gradfacts()
        return self.getGradParams(X, covGrad) * self._gradfacts()

    def getGradParams2(self, X, X2, covGrad2, want_gX=False):
        """CKern::getGradParams(g, X, X2, covGrad) (CKern.h:199-213 + overrides), natural parameters; with want_gX also
        gX[i, :] = sum_j covGrad2[i, j] d k(X_i, X2_j) / d X_i -- CKern::getGradX (CKern.h:68-74) contracted with covGrad2."""
        ctx = self._ctx_for(X)
        try:
            arr, n, keep = self._kcomps()
            X2 = fmat(X2)
            cg = fmat(covGrad2)
            g = np.zeros(self.getNumParams())
            gX = np.zeros((ctx.N, ctx.D), order="F") if want_gX else None
            check(lib().gpc_kern_grad_cross(ctx.handle, arr, n, ptr(X2), X2.shape[0], X2.shape[0], ptr(cg), cg.shape[0], ptr(g),
                                            ptr(gX) if want_gX else None))
            return (g, gX) if want_gX else g
        finally:
            ctx.close()


class CWhiteKern(CKern):
    type_code, type_name = WHITE, "white"

    def setInitParam(self):
        self.params = np.array([math.exp(-2.0)])  # CKern.cpp:641-644


class CBiasKern(CKern):
    type_code, type_name = BIAS, "bias"

    def setInitParam(self):
        self.params = np.array([math.exp(-2.0)])  # CKern.cpp:928-931


class CRbfKern(CKern):
    type_code, type_name = RBF, "rbf"

    def setInitParam(self):
        self.params = np.array([1.0, 1.0])  # inverseWidth, variance (CKern.cpp:1068-1072)


class CRbfardKern(CKern):
    type_code, type_name = RBFARD, "rbfard"

    def setInitParam(self):
        self.params = np.concatenate([[1.0, 1.0], np.full(self.inputDim, 0.5)])  # CKern.cpp:3199-3217


class CMatern32Kern(CKern):
    type_code, type_name = MATERN32, "matern32"

    def setInitParam(self):
        self.params = np.array([1.0, 1.0])  # lengthScale, variance


class CMatern52Kern(CKern):
    type_code, type_name = MATERN52, "matern52"

    def setInitParam(self):
        self.params = np.array([1.0, 1.0])


class CLinKern(CKern):
    type_code, type_name = LIN, "lin"

    def setInitParam(self):
        self.params = np.array([1.0])


class CPolyKern(CKern):
    type_code, type_name = POLY, "poly"

    def setInitParam(self):
        self.params = np.array([1.0, 1.0, 1.0])  # weightVariance, biasVariance, variance; degree 2 (CKern.cpp:2723-2729)
        self.degree = 2.0

    def setDegree(self, val):
        self.degree = float(val)

    def getDegree(self):
        return self.degree


KERN_CLASSES = {"white": CWhiteKern, "bias": CBiasKern, "rbf": CRbfKern, "rbfard": CRbfardKern,
                "matern32": CMatern32Kern, "matern52": CMatern52Kern, "lin": CLinKern, "poly": CPolyKern}


class CCmpndKern(CKern):
    """Sum of components (CKern.h:475-517, CKern.cpp:128-328); addKern clones (CKern.h:382-391)."""
    type_name = "cmpnd"

    def __init__(self, inDim_or_X):
        self.components = []
        self.inputDim = int(inDim_or_X if np.isscalar(inDim_or_X) else np.asarray(inDim_or_X).shape[1])
        self.degree = 2.0

    def setInitParam(self):
        for k in self.components:
            k.setInitParam()

    def addKern(self, kern):
        import copy
        self.components.append(copy.deepcopy(kern))
        return len(self.components) - 1

    def getNumParams(self):
        return sum(k.getNumParams() for k in self.components)

    def _components(self):
        return self.components

    def _locate(self, i):
        for k in self.components:
            n = k.getNumParams()
            if i < n:
                return k, i
            i -= n
        raise IndexError("Requested parameter doesn't exist.")

    def setParam(self, val, paramNo):
        k, i = self._locate(paramNo)
        k.setParam(val, i)

    def getParam(self, paramNo):
        k, i = self._locate(paramNo)
        return k.getParam(i)

    @property
    def params(self):
        return np.concatenate([k.params for k in self.components])

    def setParams(self, p):
        p = np.asarray(p, dtype=np.float64).ravel()
        pos = 0
        for k in self.components:
            n = k.getNumParams()
            k.setParams(p[pos:pos + n])
            pos += n

    def _transform(self, i):
        k, j = self._locate(i)
        return k._transform(j)

    def setTransParam(self, val, i):
        k, j = self._locate(i)
        k.setTransParam(val, j)


def make_kern(types, D, tparams=None):
    """cmpnd(types...) as gp.cpp:240-349 assembles it, optionally at transformed parameters."""
    kern = CCmpndKern(D)
    for t in types:
        kern.addKern(KERN_CLASSES[t](D))
    if tparams is not None:
        kern.setTransParams(tparams)
    return kern
