"""oracle/refbind.py -- ctypes binding to oracle/_ref/libgpcref.so (the UNMODIFIED reference compiled by
oracle/build_ref.sh + the oracle/ref_driver.cpp facade).  TEST INFRASTRUCTURE ONLY: importable from tests/,
__graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference); never from gpc_b200/.
"""
import ctypes as C
import os

import numpy as np

from .gp_oracle import KERN_TYPES, nparams

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libgpcref.so"))


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(os.path.join(_HERE, "_ref", "libgpcref.so"))
        _LIB.ref_last_error.restype = C.c_char_p
    return _LIB


def _d(a):
    return a.ctypes.data_as(_dp)


def _f(a):
    """column-major fp64 copy (CMatrix layout, CMatrix.h:268)."""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _spec(types, tparams):
    t = np.array([KERN_TYPES[x] for x in types], dtype=np.int32)
    p = np.ascontiguousarray(np.asarray(tparams, dtype=np.float64))
    return t, p


def _chk(rc):
    if rc < 0:
        raise RuntimeError("reference: " + lib().ref_last_error().decode())
    return rc


def set_threads(n):
    lib().ref_set_threads(int(n))


def get_threads():
    return int(lib().ref_get_threads())


def kern_params(types, tparams, D):
    t, p = _spec(types, tparams)
    out = np.zeros(len(p))
    _chk(lib().ref_kern_params(len(t), t.ctypes.data_as(_ip), _d(p), D, _d(out)))
    return out


def kern_compute(types, tparams, X):
    X = _f(X)
    N, D = X.shape
    t, p = _spec(types, tparams)
    K = np.zeros((N, N), order="F")
    _chk(lib().ref_kern_compute(len(t), t.ctypes.data_as(_ip), _d(p), _d(X), N, D, _d(K)))
    return K


def kern_cross(types, tparams, X, X2):
    X, X2 = _f(X), _f(X2)
    N, D = X.shape
    N2 = X2.shape[0]
    t, p = _spec(types, tparams)
    K = np.zeros((N, N2), order="F")
    _chk(lib().ref_kern_cross(len(t), t.ctypes.data_as(_ip), _d(p), _d(X), N, _d(X2), N2, D, _d(K)))
    return K


def kern_diag(types, tparams, X):
    X = _f(X)
    N, D = X.shape
    t, p = _spec(types, tparams)
    d = np.zeros(N)
    _chk(lib().ref_kern_diag(len(t), t.ctypes.data_as(_ip), _d(p), _d(X), N, D, _d(d)))
    return d


def kern_grad(types, tparams, X, covGrad, X2=None):
    X = _f(X)
    N, D = X.shape
    t, p = _spec(types, tparams)
    g = np.zeros(len(p))
    cg = _f(covGrad)
    if X2 is None:
        _chk(lib().ref_kern_grad(len(t), t.ctypes.data_as(_ip), _d(p), _d(X), N, D, _d(cg), _d(g)))
    else:
        X2 = _f(X2)
        _chk(lib().ref_kern_grad2(len(t), t.ctypes.data_as(_ip), _d(p), _d(X), N, _d(X2), X2.shape[0], D, _d(cg), _d(g)))
    return g


def kern_gradX(types, tparams, X, X2):
    """out[i, k, j] = d k(X_i, X2_k)/d X_ij."""
    X, X2 = _f(X), _f(X2)
    N, D = X.shape
    N2 = X2.shape[0]
    t, p = _spec(types, tparams)
    out = np.zeros((N, D, N2))  # N blocks of (N2 x D) column-major
    _chk(lib().ref_kern_gradX(len(t), t.ctypes.data_as(_ip), _d(p), _d(X), N, _d(X2), N2, D, _d(out)))
    return out.transpose(0, 2, 1).copy()


def kern_diagGradX(types, tparams, X):
    X = _f(X)
    N, D = X.shape
    t, p = _spec(types, tparams)
    out = np.zeros((N, D), order="F")
    _chk(lib().ref_kern_diagGradX(len(t), t.ctypes.data_as(_ip), _d(p), _d(X), N, D, _d(out)))
    return out


def chol_upper(A):
    A = _f(A)
    n = A.shape[0]
    U = np.zeros((n, n), order="F")
    rc = _chk(lib().ref_chol(_d(A), n, _d(U)))
    return U, rc


def jitchol(A):
    A = _f(A).copy(order="F")
    n = A.shape[0]
    U = np.zeros((n, n), order="F")
    jit = C.c_double(0)
    _chk(lib().ref_jitchol(_d(A), n, _d(U), C.byref(jit)))
    return U, jit.value, A


def pdinv(U):
    U = _f(U)
    n = U.shape[0]
    inv = np.zeros((n, n), order="F")
    ld = C.c_double(0)
    _chk(lib().ref_pdinv(_d(U), n, _d(inv), C.byref(ld)))
    return inv, ld.value


def trsm(B, T, alpha, side, uplo, trans, diag):
    B = _f(B).copy(order="F")
    T = _f(T)
    m, n = B.shape
    _chk(lib().ref_trsm(_d(B), m, n, _d(T), C.c_double(alpha), side.encode(), uplo.encode(), trans.encode(),
                        diag.encode()))
    return B


def syrk(Cm, A, alpha, beta, uplo, trans):
    Cm = _f(Cm).copy(order="F")
    A = _f(A)
    _chk(lib().ref_syrk(_d(Cm), Cm.shape[0], _d(A), A.shape[0], A.shape[1], C.c_double(alpha), C.c_double(beta),
                        uplo.encode(), trans.encode()))
    return Cm


def gemm(Cm, A, B, alpha, beta, ta, tb):
    Cm = _f(Cm).copy(order="F")
    A, B = _f(A), _f(B)
    _chk(lib().ref_gemm(_d(Cm), Cm.shape[0], Cm.shape[1], _d(A), A.shape[0], A.shape[1], _d(B), B.shape[0],
                        B.shape[1], C.c_double(alpha), C.c_double(beta), ta.encode(), tb.encode()))
    return Cm


def gp_eval(types, tparams, X, y, bias=None, scale=None, Xs=None, reps=1):
    """CGp FTC evaluation through the reference.  Returns dict(ll, g, mu, var, t_ll, t_grad, t_eval)."""
    X = _f(X)
    N, D = X.shape
    y = _f(np.asarray(y, dtype=np.float64).reshape(N, -1))
    d = y.shape[1]
    bias = np.zeros(d) if bias is None else np.ascontiguousarray(np.asarray(bias, dtype=np.float64).reshape(d))
    scale = np.ones(d) if scale is None else np.ascontiguousarray(np.asarray(scale, dtype=np.float64).reshape(d))
    t, p = _spec(types, tparams)
    ll = C.c_double(0)
    g = np.zeros(len(p) + 8)
    tim = np.zeros(3)
    mu = var = None
    if Xs is not None:
        Xs = _f(Xs)
        Ns = Xs.shape[0]
        mu = np.zeros((Ns, d), order="F")
        var = np.zeros((Ns, d), order="F")
        xs_p, mu_p, var_p = _d(Xs), _d(mu), _d(var)
    else:
        Ns, xs_p, mu_p, var_p = 0, None, None, None
    P = _chk(lib().ref_gp_eval(len(t), t.ctypes.data_as(_ip), _d(p), _d(X), _d(y), N, D, d, _d(bias), _d(scale),
                               C.byref(ll), _d(g), xs_p, Ns, mu_p, var_p, _d(tim), int(reps)))
    return dict(ll=ll.value, g=g[:P].copy(), mu=mu, var=var, t_ll=tim[0], t_grad=tim[1], t_eval=tim[2])


def phase_times(types, tparams, X):
    X = _f(X)
    N, D = X.shape
    t, p = _spec(types, tparams)
    out = np.zeros(4)
    _chk(lib().ref_phase_times(len(t), t.ctypes.data_as(_ip), _d(p), _d(X), N, D, _d(out)))
    return dict(kern_compute=out[0], dpotrf=out[1], pdinv=out[2], trans=out[3])


def gplvm_eval(types, tparams, Xlat, Y):
    Xlat, Y = _f(Xlat), _f(Y)
    N, q = Xlat.shape
    d = Y.shape[1]
    t, p = _spec(types, tparams)
    ll = C.c_double(0)
    g = np.zeros(len(p) + N * q + d + 8)
    m = np.zeros((N, d), order="F")
    P = _chk(lib().ref_gplvm_eval(len(t), t.ctypes.data_as(_ip), _d(p), _d(Xlat), _d(Y), N, q, d, C.byref(ll),
                                  _d(g), _d(m)))
    return dict(ll=ll.value, g=g[:P].copy(), m=m)


def gplvm_initX(Y, q):
    Y = _f(Y)
    N, d = Y.shape
    X = np.zeros((N, q), order="F")
    _chk(lib().ref_gplvm_initX(_d(Y), N, q, d, _d(X)))
    return X


def svml_read(path):
    """CClctrl::readSvmlDataFile (CClctrl.cpp:55-171) of the compiled reference: (X, y)."""
    L = lib()
    n, d = C.c_int(0), C.c_int(0)
    _chk(L.ref_svml_read(path.encode(), C.byref(n), C.byref(d), None, None))
    X = np.zeros((n.value, d.value), order="F")
    y = np.zeros((n.value, 1), order="F")
    _chk(L.ref_svml_read(path.encode(), C.byref(n), C.byref(d), _d(X), _d(y)))
    return X, y


def gp_optimise(types, tparams, X, y, iters, bias=None, scale=None):
    """CGp::optimise with SCG (gp learn's default) for `iters` iterations: (transformed parameters, log-likelihood)."""
    L = lib()
    X, y = _f(X), _f(y)
    N, D = X.shape
    dout = y.shape[1]
    codes, tp = _spec(types, tparams)
    bias = np.zeros(dout) if bias is None else np.asarray(bias, dtype=np.float64).ravel()
    scale = np.ones(dout) if scale is None else np.asarray(scale, dtype=np.float64).ravel()
    out = np.zeros(tp.size)
    ll = C.c_double(0)
    _chk(L.ref_gp_optimise(len(codes), codes.ctypes.data_as(_ip), _d(tp), _d(X), _d(y), N, D, dout,
                           _d(bias), _d(scale), int(iters), _d(out), C.byref(ll)))
    return out, ll.value
