"""oracle/gp_oracle.py -- CPU restatement (numpy, fp64) of GPc's exact-GP hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in gpc_b200/ may import this module: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and only as
the checker.  The product path (gpc_b200 -> libgpc_b200.so) fails loudly without its CUDA library.

PARITY IS PINNED: every function here is checked (tests/test_oracle_cpu.py) against
  * the reference's own MATLAB-generated golden vectors (matfiles/*KernTest.mat, testGpftc.mat,
    choleskyMatrixTest.mat ... committed as tests/golden/*.npz by tests/golden/make_golden.py), and
  * outputs of the unmodified reference compiled here (oracle/_ref/libgpcref.so, oracle/refbind.py).

All citations are file:line into the reference (SheffieldML/GPc @ 01242d9f).
Matrices are numpy fp64; X is N x D.  A kernel is a list of components
    [(type_name, params), ...]      params = UNTRANSFORMED values in the reference's order
forming CCmpndKern's sum (CKern.cpp:128-328).
"""
import math

import numpy as np

KERN_TYPES = {"white": 0, "bias": 1, "rbf": 2, "rbfard": 3, "matern32": 4, "matern52": 5, "lin": 6, "poly": 7}
KERN_NAMES = {v: k for k, v in KERN_TYPES.items()}
LIMVAL = 36.0  # CTransform.h:17
EPS = np.finfo(np.float64).eps  # ndlutil.h:34
HALFLOGTWOPI = 0.5 * math.log(2.0 * math.pi)  # ndlutil.h:39
POLY_DEGREE = 2.0  # CKern.cpp:2723-2729 (setInitParam: degree = 2, not an optimised parameter)


def nparams(ktype, D):
    """Parameter count per component: white/bias/lin 1, rbf/matern 2, poly 3, rbfard 2+D
    (CKern.cpp:636, 921, 1059, 1734, 1981, 2252, 2707, 3189-3217)."""
    return {"white": 1, "bias": 1, "lin": 1, "rbf": 2, "matern32": 2, "matern52": 2, "poly": 3, "rbfard": 2 + D}[ktype]


def transform_kinds(ktype, D):
    """'exp' (defaultPositive) for every parameter except rbfard's input scales, which are
    'sigmoid' (defaultZeroOne): CKern.cpp:1062-1065, 3193-3217."""
    n = nparams(ktype, D)
    if ktype == "rbfard":
        return ["exp", "exp"] + ["sigmoid"] * D
    return ["exp"] * n


def atox(a, kind):
    """transformed -> natural parameter.  CExpTransform::atox CTransform.cpp:31-42 (exponent clamped
    to +-36); CSigmoidTransform::atox CTransform.cpp:96-104 (clamped to [eps, 1-eps])."""
    if kind == "exp":
        return math.exp(min(max(a, -LIMVAL), LIMVAL))
    if kind == "sigmoid":
        if a < -LIMVAL:
            return EPS
        if a < LIMVAL:
            return 1.0 / (1.0 + math.exp(-a))
        return 1.0 - EPS
    raise ValueError(kind)


def xtoa(x, kind):
    """CExpTransform::xtoa CTransform.cpp:43-49; CSigmoidTransform::xtoa :105-108."""
    if kind == "exp":
        return math.log(x)
    if kind == "sigmoid":
        return math.log(x / (1.0 - x))
    raise ValueError(kind)


def gradfact(x, kind):
    """d x / d a expressed in x.  exp: x (CTransform.cpp:50-53); sigmoid: x(1-x) (:109-112)."""
    return x if kind == "exp" else x * (1.0 - x)


def kern_from_trans(types, tparams, D):
    """Split a compound transformed-parameter vector (CCmpndKern order, CKern.h:382-391) into
    [(type, natural params)]."""
    out, pos = [], 0
    for t in types:
        kinds = transform_kinds(t, D)
        p = np.array([atox(float(tparams[pos + i]), kinds[i]) for i in range(len(kinds))])
        out.append((t, p))
        pos += len(kinds)
    assert pos == len(tparams)
    return out


def trans_from_kern(kern, D):
    out = []
    for t, p in kern:
        kinds = transform_kinds(t, D)
        out += [xtoa(float(p[i]), kinds[i]) for i in range(len(kinds))]
    return np.array(out)


def _dist2(X, X2):
    """CMatrix::dist2Row CMatrix.h:554-560: |x|^2 + |y|^2 - 2 x.y (NOT direct differences)."""
    n1 = np.sum(X * X, axis=1)[:, None]
    n2 = np.sum(X2 * X2, axis=1)[None, :]
    return n1 + n2 - 2.0 * (X @ X2.T)


def _clamp0(d2):
    """The reference takes sqrt(dist2Row) unguarded (CKern.cpp:1839, 2093): rounding can make r^2 slightly
    negative for (near-)identical rows and the result NaN.  Only diagonal entries hit this in the paths we
    check (they are overwritten by diagComputeElement), so the oracle clamps at 0 -- as the CUDA path does."""
    return np.maximum(d2, 0.0)


def _ard_dist2(X, X2, s):
    """CRbfardKern::computeElement CKern.cpp:3305-3316: sum_k s_k (x_ik - x_jk)^2, direct differences."""
    d = X[:, None, :] - X2[None, :, :]
    return np.einsum("ijk,k->ij", d * d, s)


def _cross_component(t, p, X, X2):
    """computeElement for every (i, j): CKern.cpp:702-706 white (always 0), :989-993 bias, :1147-1154 rbf,
    :3305-3316 rbfard, :1834-1842 matern32, :2087-2096 matern52, :2328-2332 lin, :2815-2820 poly."""
    if t == "white":
        return np.zeros((X.shape[0], X2.shape[0]))
    if t == "bias":
        return np.full((X.shape[0], X2.shape[0]), p[0])
    if t == "rbf":
        return p[1] * np.exp(-0.5 * p[0] * _dist2(X, X2))
    if t == "rbfard":
        return p[1] * np.exp(-0.5 * p[0] * _ard_dist2(X, X2, p[2:]))
    if t == "matern32":
        z = np.sqrt(_clamp0(_dist2(X, X2)) * (3.0 / (p[0] * p[0])))
        return p[1] * (1.0 + z) * np.exp(-z)
    if t == "matern52":
        zz = _clamp0(_dist2(X, X2)) * (5.0 / (p[0] * p[0]))
        z = np.sqrt(zz)
        return p[1] * (1.0 + z + zz / 3.0) * np.exp(-z)
    if t == "lin":
        return p[0] * (X @ X2.T)
    if t == "poly":
        return p[2] * np.power(p[0] * (X @ X2.T) + p[1], POLY_DEGREE)
    raise ValueError(t)


def _diag_component(t, p, X):
    """diagComputeElement: white CKern.cpp:646-649, bias :933-936, rbf :1074-1077, rbfard :3219-3222,
    matern :1762, :2009 (variance), lin :2262-2265, poly :2731-2735."""
    n = X.shape[0]
    if t in ("white", "bias", "lin"):
        if t == "lin":
            return p[0] * np.sum(X * X, axis=1)
        return np.full(n, p[0])
    if t in ("rbf", "rbfard", "matern32", "matern52"):
        return np.full(n, p[1])
    if t == "poly":
        return p[2] * np.power(p[0] * np.sum(X * X, axis=1) + p[1], POLY_DEGREE)
    raise ValueError(t)


def kern_diag(kern, X):
    """CCmpndKern::diagComputeElement CKern.cpp:165-171."""
    return sum(_diag_component(t, p, X) for t, p in kern)


def kern_cross(kern, X, X2):
    """CKern::compute(K, X, X2) CKern.h:146-157 over CCmpndKern::computeElement CKern.cpp:219-226."""
    return sum(_cross_component(t, p, X, X2) for t, p in kern)


def kern_compute(kern, X):
    """CKern::compute(K, X) CKern.h:128-144 == CGp::_updateK FTC CGp.cpp:693-712:
    off-diagonal from computeElement, diagonal from diagComputeElement (this is where white enters)."""
    K = kern_cross(kern, X, X)
    K = 0.5 * (K + K.T)
    np.fill_diagonal(K, kern_diag(kern, X))
    return K


def _grad_component(t, p, X, X2, cg, sym):
    """getGradParams for one component: sum_ij covGrad_ij dK_ij/dtheta (natural parameters).
    sym=True: the (X, covGrad) overloads where the diagonal follows diagComputeElement semantics
    (rbf CKern.cpp:1204-1241, rbfard :3359-3403, matern32 :1895-1938, matern52 :2156-2202, lin :2369-2383,
    poly :2848-2891, white :735-739 (trace), bias :1020-1024 (sum)); sym=False: the (X, X2, covGrad)
    overloads (rbf :1175-1202, white -> 0, ...)."""
    if t == "white":
        return np.array([np.trace(cg) if sym else 0.0])
    if t == "bias":
        return np.array([cg.sum()])
    if t == "rbf":
        d2 = _dist2(X, X2)
        if sym:
            np.fill_diagonal(d2, 0.0)
        k = np.exp(-0.5 * p[0] * d2)
        return np.array([-0.5 * p[1] * np.sum(d2 * k * cg), np.sum(k * cg)])
    if t == "rbfard":
        s = p[2:]
        d = X[:, None, :] - X2[None, :, :]
        dd = d * d
        val = np.einsum("ijk,k->ij", dd, s)
        kcg = np.exp(-0.5 * p[0] * val) * cg
        g1 = -0.5 * p[1] * np.sum(val * kcg)
        g2 = np.sum(kcg)
        gs = -0.5 * p[0] * p[1] * np.einsum("ij,ijk->k", kcg, dd)
        return np.concatenate([[g1, g2], gs])
    if t in ("matern32", "matern52"):
        c = 3.0 if t == "matern32" else 5.0
        d2 = _clamp0(_dist2(X, X2))
        if sym:
            np.fill_diagonal(d2, 0.0)
        zz = d2 * (c / (p[0] * p[0]))
        z = np.sqrt(zz)
        e = np.exp(-z)
        if t == "matern32":
            k = (1.0 + z) * e
            dl = zz * e / p[0]  # (wi2/l)(n2/z)(k - e) = z^2 e^-z / l
        else:
            k = (1.0 + z + zz / 3.0) * e
            dl = (zz / 3.0) * (1.0 + z) * e / p[0]
        return np.array([p[1] * np.sum(cg * dl), np.sum(cg * k)])
    if t == "lin":
        return np.array([np.sum(cg * (X @ X2.T))])
    if t == "poly":
        ip = X @ X2.T
        arg = p[0] * ip + p[1]
        base = p[2] * POLY_DEGREE * np.power(arg, POLY_DEGREE - 1.0) * cg
        return np.array([np.sum(ip * base), np.sum(base), np.sum(np.power(arg, POLY_DEGREE) * cg)])
    raise ValueError(t)


def kern_grad_params(kern, X, covGrad, X2=None):
    """CCmpndKern::getGradParams CKern.cpp:284-298 (natural-parameter gradients, component order)."""
    sym = X2 is None
    X2 = X if sym else X2
    return np.concatenate([_grad_component(t, p, X, X2, covGrad, sym) for t, p in kern])


def kern_grad_trans_params(kern, X, covGrad, X2=None):
    """CKern::getGradTransParams CKern.cpp:36-63: natural gradient x gradfact(theta)."""
    g = kern_grad_params(kern, X, covGrad, X2)
    D = X.shape[1]
    pos = 0
    for t, p in kern:
        kinds = transform_kinds(t, D)
        for i, kd in enumerate(kinds):
            g[pos + i] *= gradfact(float(p[i]), kd)
        pos += len(kinds)
    return g


def kern_gradX(kern, X, X2):
    """CKern::getGradX(gX, X, X2) CKern.h:68-74: out[i, k, j] = d k(X_i, X2_k) / d X_{i,j}  (note: the
    derivative is wrt the FIRST argument's row; sign convention pf*(x2 - x)).
    rbf CKern.cpp:1115-1135, rbfard :3268-3293, matern32 :1796-1822, matern52 :2042-2075, lin :2291-2308,
    poly :2774-2792, white/bias -> 0."""
    N, D = X.shape
    N2 = X2.shape[0]
    out = np.zeros((N, N2, D))
    diff = X2[None, :, :] - X[:, None, :]  # (x2_k - x_i)
    for t, p in kern:
        if t in ("white", "bias"):
            continue
        if t == "rbf":
            k = np.exp(-0.5 * p[0] * _dist2(X, X2))
            out += (p[1] * p[0]) * diff * k[:, :, None]
        elif t == "rbfard":
            s = p[2:]
            k = np.exp(-0.5 * p[0] * _ard_dist2(X, X2, s))
            out += (p[1] * p[0]) * diff * k[:, :, None] * s[None, None, :]
        elif t == "matern32":
            wi2 = 3.0 / (p[0] * p[0])
            z = np.sqrt(_clamp0(_dist2(X, X2)) * wi2)
            out += (p[1] * wi2) * diff * np.exp(-z)[:, :, None]
        elif t == "matern52":
            wi2 = 5.0 / (p[0] * p[0])
            z = np.sqrt(_clamp0(_dist2(X, X2)) * wi2)
            out += (p[1] * wi2 / 3.0) * diff * ((1.0 + z) * np.exp(-z))[:, :, None]
        elif t == "lin":
            out += p[0] * X2[None, :, :]
        elif t == "poly":
            arg = p[0] * (X @ X2.T) + p[1]
            kv = POLY_DEGREE * p[2] * p[0] * np.power(arg, POLY_DEGREE - 1.0)
            out += kv[:, :, None] * X2[None, :, :]
    return out


def kern_diagGradX(kern, X):
    """getDiagGradX: zero for stationary kernels; lin 2 v x (CKern.cpp:2310-2322); poly :2793-2809."""
    out = np.zeros_like(X)
    for t, p in kern:
        if t == "lin":
            out += 2.0 * p[0] * X
        elif t == "poly":
            arg = p[0] * np.sum(X * X, axis=1) + p[1]
            kv = POLY_DEGREE * p[2] * p[0] * np.power(arg, POLY_DEGREE - 1.0)
            out += 2.0 * kv[:, None] * X
    return out


# ---------------------------------------------------------------------------------------------
# Dense primitives (CMatrix.cpp) -- plain loops / numpy so that they do not depend on LAPACK.
def chol_lower(A):
    """Lower Cholesky (CMatrix::chol CMatrix.cpp:380-403 computes the upper U = L^T through dpotrf_,
    lapack.h:59-65).  Returns (L, info): info = 1-based order of the first non-positive pivot, 0 if PD."""
    A = np.array(A, dtype=np.float64)
    n = A.shape[0]
    L = np.zeros_like(A)
    for j in range(n):
        v = A[j, j] - np.dot(L[j, :j], L[j, :j])
        if not (v > 0.0):
            return L, j + 1
        L[j, j] = math.sqrt(v)
        if j + 1 < n:
            L[j + 1:, j] = (A[j + 1:, j] - L[j + 1:, :j] @ L[j, :j]) / L[j, j]
    return L, 0


def jit_chol(A, max_tries=20):
    """CMatrix::jitChol CMatrix.cpp:767-804: jitter0 = 1e-6 * trace(A)/n; on MatrixNonPosDef the diagonal of A
    (the CALLER's matrix: it is mutated) gets += jitter, then jitter *= 10, and it throws once jitter > 10 or
    after max_tries.  The value returned is the running `jitter` variable: jitter0 if the first try succeeds,
    otherwise 10x the last amount added."""
    A = np.array(A, dtype=np.float64)
    n = A.shape[0]
    jitter = 1e-6 * np.trace(A) / n
    tries = 0
    while tries < max_tries:
        L, info = chol_lower(A)
        if info == 0:
            return L, jitter, A
        A[np.diag_indices(n)] += jitter
        jitter *= 10.0
        tries += 1
        if jitter > 10.0:
            break
    raise np.linalg.LinAlgError("matrix is non positive definite")


def log_det(L):
    """logDet CMatrix.cpp:404-412: 2 sum log U_ii."""
    return 2.0 * float(np.sum(np.log(np.diag(L))))


def pdinv(L):
    """CMatrix::pdinv CMatrix.cpp:421-432 (dpotri_, lapack.h:67-73): K^-1 = L^-T L^-1, full symmetric."""
    n = L.shape[0]
    Li = solve_lower(L, np.eye(n))
    Kinv = Li.T @ Li
    return 0.5 * (Kinv + Kinv.T)


def solve_lower(L, B, trans=False):
    """dtrsm_ 'L','L',trans,'N' (CMatrix.cpp:272-295): forward / backward substitution, blocked by rows."""
    B = np.array(B, dtype=np.float64)
    n = L.shape[0]
    X = B.copy()
    if not trans:
        for i in range(n):
            X[i] = (X[i] - L[i, :i] @ X[:i]) / L[i, i]
    else:
        for i in range(n - 1, -1, -1):
            X[i] = (X[i] - L[i + 1:, i] @ X[i + 1:]) / L[i, i]
    return X


# ---------------------------------------------------------------------------------------------
# CGp FTC path
def gp_loglik_grad(kern, X, y, bias=None, scale=None):
    """One CGp evaluation (FTC, spherical):
      m = (y - bias)/scale                                      CGp::updateM      CGp.cpp:248-260
      K                                                         CGp::_updateK     CGp.cpp:693-712
      L = jitChol(K), logdet, K^-1                              CGp::_updateInvK  CGp.cpp:877-891
      ll = -1/2 sum_j (m_j' K^-1 m_j + logdet) - d N/2 log 2pi  CGp::logLikelihood CGp.cpp:913-938,1002-1013
      covGrad_j = -1/2 (K^-1 - a_j a_j'),  a_j = K^-1 m_j       CGp::updateCovGradient CGp.cpp:666-679
      g = sum_j getGradTransParams(X, covGrad_j)                CGp::updateG      CGp.cpp:1096-1116
    Returns dict(ll, g (transformed-parameter gradient), K, L, Kinv, alpha, logdet, m)."""
    y = np.asarray(y, dtype=np.float64).reshape(X.shape[0], -1)
    N, d = y.shape
    bias = np.zeros(d) if bias is None else np.asarray(bias, dtype=np.float64).reshape(d)
    scale = np.ones(d) if scale is None else np.asarray(scale, dtype=np.float64).reshape(d)
    m = (y - bias[None, :]) / scale[None, :]
    K = kern_compute(kern, X)
    L, _, K = jit_chol(K)
    logdet = log_det(L)
    Kinv = pdinv(L)
    alpha = Kinv @ m
    ll = -0.5 * (float(np.sum(alpha * m)) + d * logdet) - d * N * HALFLOGTWOPI
    g = np.zeros(sum(nparams(t, X.shape[1]) for t, _ in kern))
    for j in range(d):
        cg = -0.5 * (Kinv - np.outer(alpha[:, j], alpha[:, j]))
        g += kern_grad_trans_params(kern, X, cg)
    return dict(ll=ll, g=g, K=K, L=L, Kinv=Kinv, alpha=alpha, logdet=logdet, m=m)


def gp_posterior(kern, X, y, Xs, bias=None, scale=None):
    """CGp::posteriorMeanVar CGp.cpp:642-663: mu = K*' alpha * scale + bias (:548-574);
    var = (k(x*,x*) - |L^-1 k*|^2) * scale^2 (:600-623); alpha by two triangular solves (:469-484)."""
    y = np.asarray(y, dtype=np.float64).reshape(X.shape[0], -1)
    N, d = y.shape
    bias = np.zeros(d) if bias is None else np.asarray(bias, dtype=np.float64).reshape(d)
    scale = np.ones(d) if scale is None else np.asarray(scale, dtype=np.float64).reshape(d)
    m = (y - bias[None, :]) / scale[None, :]
    K = kern_compute(kern, X)
    L, _, K = jit_chol(K)
    alpha = solve_lower(L, solve_lower(L, m), trans=True)
    kX = kern_cross(kern, X, Xs)
    mu = (kX.T @ alpha) * scale[None, :] + bias[None, :]
    V = solve_lower(L, kX)
    v = kern_diag(kern, Xs) - np.sum(V * V, axis=0)
    var = v[:, None] * (scale * scale)[None, :]
    return mu, var


def gplvm_loglik_grad(kern, X, m):
    """CGplvm (plain FTC, latent-regularised, no dynamics / back-constraints):
      ll = -1/2 [ sum_j (m_j' K^-1 m_j + logdet) + sum_k |X_:k|^2 ]   CGplvm::logLikelihood CGplvm.cpp:493-553
           (no -dN/2 log 2pi term, unlike CGp)
      gX_ik = sum_j' covGrad_i,j' * 2 dk_ij'/dx_ik (diagonal row replaced by getDiagGradX) - X_ik
                                                                      CGplvm.cpp:569-603, 672-681
      g = [kernel trans-param gradients][gX col-major]                CGplvm.cpp:257-290
    CGplvm::_updateInvK uses chol() without jitter (CGplvm.cpp:435-446)."""
    N, q = X.shape
    d = m.shape[1]
    K = kern_compute(kern, X)
    L, info = chol_lower(K)
    if info:
        raise np.linalg.LinAlgError("matrix is non positive definite")
    logdet = log_det(L)
    Kinv = pdinv(L)
    alpha = Kinv @ m
    ll = -0.5 * (float(np.sum(alpha * m)) + d * logdet + float(np.sum(X * X)))
    gk = np.zeros(sum(nparams(t, q) for t, _ in kern))
    cgsum = np.zeros((N, N))
    for j in range(d):
        cg = -0.5 * (Kinv - np.outer(alpha[:, j], alpha[:, j]))
        gk += kern_grad_trans_params(kern, X, cg)
        cgsum += cg
    G = 2.0 * kern_gradX(kern, X, X)  # [i, k, j]
    dg = kern_diagGradX(kern, X)
    for i in range(N):
        G[i, i, :] = dg[i, :]
    gX = np.einsum("ikj,ki->ij", G, cgsum) - X
    return dict(ll=ll, g=np.concatenate([gk, gX.T.reshape(-1)]), gk=gk, gX=gX)
