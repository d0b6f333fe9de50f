"""CMatrix-level mirror: the dense methods on the hot path (CMatrix.h:1055-1113, CMatrix.cpp) as functions over
column-major numpy arrays, executed by libgpc_b200.so (gpc_dpotrf / gpc_dpotri / gpc_dtrsm / gpc_dsyrk / gpc_dgemm /
gpc_dsymv -- the drop-ins for the lapack.h symbols CMatrix calls)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, fmat, lib, ptr


def _c(ch):
    return C.c_char(ch.encode()[0:1])


def potrf(A, uplo="U", device=0):
    """CMatrix::potrf (CMatrix.cpp:371-379): in-place dpotrf_; raises MatrixNonPosDef when info != 0."""
    A = fmat(A).copy(order="F")
    info = C.c_int(0)
    rc = check(lib().gpc_dpotrf(device, _c(uplo), A.shape[0], ptr(A), A.shape[0], C.byref(info)))
    if rc > 0:
        raise _lib.MatrixNonPosDef(rc)
    return A


def chol(A, uplo="U", device=0):
    """CMatrix::chol (CMatrix.cpp:380-403): potrf + zero the other triangle."""
    A = potrf(A, uplo, device)
    return np.triu(A) if uplo.upper() == "U" else np.tril(A)


def jitChol(A, maxTries=20, device=0):
    """CMatrix::jitChol (CMatrix.cpp:767-804).  Returns (U, jitter, A_mutated)."""
    A = fmat(A).copy(order="F")
    n = A.shape[0]
    jitter = 1e-6 * np.trace(A) / n
    tries = 0
    while tries < maxTries:
        try:
            return chol(A, "U", device), jitter, A
        except _lib.MatrixNonPosDef:
            A[np.diag_indices(n)] += jitter
            jitter *= 10.0
            tries += 1
            if jitter > 10.0:
                raise
    raise _lib.MatrixNonPosDef(1)


def potri(U, uplo="U", device=0):
    """CMatrix::potri (CMatrix.cpp:414-420): dpotri_ on a factor; only the uplo triangle is meaningful."""
    A = fmat(U).copy(order="F")
    info = C.c_int(0)
    check(lib().gpc_dpotri(device, _c(uplo), A.shape[0], ptr(A), A.shape[0], C.byref(info)))
    return A


def pdinv(U, device=0):
    """CMatrix::pdinv(U) (CMatrix.cpp:421-432): potri('U') + mirror to the lower triangle."""
    A = potri(U, "U", device)
    iu = np.triu_indices(A.shape[0], 1)
    A.T[iu] = A[iu]
    return A


def logDet(U):
    """logDet (CMatrix.cpp:404-412)."""
    return 2.0 * float(np.sum(np.log(np.diag(U))))


def trsm(B, A, alpha, side, uplo, trans, diag, device=0):
    """CMatrix::trsm (CMatrix.cpp:272-295): B := alpha * op(A^-1) B or alpha * B op(A^-1)."""
    B = fmat(B).copy(order="F")
    A = fmat(A)
    m, n = B.shape
    check(lib().gpc_dtrsm(device, _c(side), _c(uplo), _c(trans), _c(diag), m, n, float(alpha), ptr(A), A.shape[0],
                          ptr(B), m))
    return B


def syrk(Cm, A, alpha, beta, uplo, trans, device=0):
    """CMatrix::syrk (CMatrix.cpp:297-322)."""
    Cm = fmat(Cm).copy(order="F")
    A = fmat(A)
    n = Cm.shape[0]
    k = A.shape[1] if trans.lower() == "n" else A.shape[0]
    check(lib().gpc_dsyrk(device, _c(uplo), _c(trans), n, k, float(alpha), ptr(A), A.shape[0], float(beta), ptr(Cm), n))
    return Cm


def gemm(Cm, A, B, alpha, beta, transa, transb, device=0):
    """CMatrix::gemm (CMatrix.cpp:205-247)."""
    Cm = fmat(Cm).copy(order="F")
    A, B = fmat(A), fmat(B)
    m, n = Cm.shape
    k = A.shape[1] if transa.lower() == "n" else A.shape[0]
    check(lib().gpc_dgemm(device, _c(transa), _c(transb), m, n, k, float(alpha), ptr(A), A.shape[0], ptr(B),
                          B.shape[0], float(beta), ptr(Cm), m))
    return Cm


def symv(y, A, x, alpha, beta, uplo, device=0):
    """CMatrix::symv (CMatrix.cpp:127-203)."""
    y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).ravel()).copy()
    A = fmat(A)
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel())
    check(lib().gpc_dsymv(device, _c(uplo), A.shape[0], float(alpha), ptr(A), A.shape[0], ptr(x), float(beta), ptr(y)))
    return y


def syr(A, x, alpha, uplo, device=0):
    """CMatrix::syr (CMatrix.h:526-533): A := alpha x x' + A on the `uplo` triangle, then mirrored (copySymmetric)."""
    A = fmat(A).copy(order="F")
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel())
    n = A.shape[0]
    check(lib().gpc_dsyr(device, _c(uplo), n, float(alpha), ptr(x), 1, ptr(A), n))
    tri = np.triu(A) if uplo.lower() == "u" else np.tril(A)
    return tri + tri.T - np.diag(np.diag(A))
