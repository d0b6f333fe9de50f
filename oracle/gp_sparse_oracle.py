"""oracle/gp_sparse_oracle.py -- TEST INFRASTRUCTURE (never imported by gpc_b200/).

Oracle for SURVEY.md 8(f) row 2, the reference's sparse approximations DTC / DTCVAR / FITC of CGp
(CGp.cpp:713-735 _updateK, :766-861 updateAD, :939-988 logLikelihood, :1244-1413 updateG): log-likelihood and the
gradient in the optimiser's parameter order [X_u column-major][kernel transformed parameters][log beta]
(CGp.cpp:330-385).  The product path for this row is NOT built yet; this file and its pins
(tests/test_oracle_sparse_cpu.py: the reference's MATLAB fixtures matfiles/testGpdtc.mat, testGpfitc.mat and the
compiled reference on seeded inputs) are what that path will be checked against.

Restatement, not transcription: all three are the Gaussian log-density of the targets under

    Sigma = Q + Lambda,  Q = K_fu K_uu^-1 K_uf,
    DTC, DTCVAR: Lambda = I/beta;        FITC: Lambda = diag(k_ii - q_ii) + I/beta,

which the reference evaluates in Woodbury form (A = K_uu/beta + K_uf K_uf', M x M factorisations); so does this file
(matrix determinant / inversion lemmas, no N x N matrix is formed), but the gradient is derived once from the density
instead of following the reference's term-by-term code: with G = dL/dSigma = -1/2 sum_j (Sigma^-1 - Sigma^-1 m_j m_j'
Sigma^-1) and B = K_uu^-1 K_uf,

    dL/dK_uf = 2 B H,   dL/dK_uu = -B H B',   dL/d diag(K) = h,   dL/dbeta = -tr(G)/beta^2
    DTC:  H = G, h = 0;   FITC: H = G - diag(G), h = diag(G);
    DTCVAR: the DTC terms plus those of the trace penalty -1/2 d beta tr(K - Q) (CGp.cpp:955-956).

Constants follow the reference exactly, including its FITC quirk: CGp.cpp:963 adds N log 2pi inside the bracket and
CGp.cpp:1012 subtracts d N/2 log 2pi again, so the FITC value is log N(.) - d N/2 log 2pi.
"""
import numpy as np

from . import gp_oracle as O

LOGTWOPI = float(np.log(2.0 * np.pi))


def _sym_gradX(kern, Xu, gKuu):
    """d sum_ab gKuu[a,b] K_uu[a,b] / d X_u: off-diagonal pairs count twice, the diagonal through getDiagGradX
    (CKern::getGradX / getDiagGradX conventions, CKern.h:68-74; used by CGp.cpp:1283-1304)."""
    G = 2.0 * O.kern_gradX(kern, Xu, Xu)  # [i, j, k] = 2 d k(xu_i, xu_j) / d xu_ik
    dg = O.kern_diagGradX(kern, Xu)
    for i in range(Xu.shape[0]):
        G[i, i, :] = dg[i, :]
    return np.einsum("ijk,ij->ik", G, gKuu)


def sparse_loglik_grad(kern, X, y, Xu, beta, approx="dtc", bias=None, scale=None):
    """returns dict(ll, g, gXu, gk, gbeta): ll as CGp::logLikelihood returns it (with -d N/2 log 2pi, CGp.cpp:1012),
    g = [gXu column-major, kernel transformed-parameter gradients, d ll / d log beta] (CGp.cpp:1016-1079)."""
    X = np.asarray(X, dtype=np.float64)
    Xu = np.asarray(Xu, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(X.shape[0], -1)
    N, D = X.shape
    M = Xu.shape[0]
    d = y.shape[1]
    bias = np.zeros(d) if bias is None else np.asarray(bias, dtype=np.float64).reshape(d)
    scale = np.ones(d) if scale is None else np.asarray(scale, dtype=np.float64).reshape(d)
    m = (y - bias[None, :]) / scale[None, :]  # CGp::updateM, CGp.cpp:248-260
    approx = approx.lower()
    if approx not in ("dtc", "dtcvar", "fitc"):
        raise ValueError("approx must be dtc, dtcvar or fitc")

    Kuu = O.kern_compute(kern, Xu)          # diagonal through diagComputeElement: white noise included (CGp.cpp:721)
    Kuf = O.kern_cross(kern, Xu, X)         # computeElement: no white noise (CGp.cpp:729-732)
    kdiag = O.kern_diag(kern, X)            # CGp.cpp:745-749 (FITC, DTCVAR)
    B = np.linalg.solve(Kuu, Kuf)           # K_uu^-1 K_uf
    qdiag = np.einsum("ij,ij->j", Kuf, B)
    lam = np.full(N, 1.0 / beta)
    if approx == "fitc":
        lam = lam + (kdiag - qdiag)
    # Everything that involves Sigma^-1 is formed through A = K_uu + K_uf Lambda^-1 K_fu (M x M) -- for DTC this is beta
    # times the reference's A (CGp.cpp:770-772) -- with the products simplified analytically so that nothing cancels:
    #   K_uu^-1 K_uf Sigma^-1 = A^-1 K_uf Lambda^-1,   B Sigma^-1 B' = K_uu^-1 - A^-1
    linv = 1.0 / lam
    KufL = Kuf * linv[None, :]
    A = Kuu + KufL @ Kuf.T
    A = 0.5 * (A + A.T)
    logdet = float(np.sum(np.log(lam))) - np.linalg.slogdet(Kuu)[1] + np.linalg.slogdet(A)[1]
    BS = np.linalg.solve(A, KufL)                                   # = B Sigma^-1 (M x N)
    a = linv[:, None] * m - KufL.T @ np.linalg.solve(A, KufL @ m)   # Sigma^-1 m (N x d)
    Ba = BS @ m                                                     # B Sigma^-1 m (M x d)
    quad = float(np.sum(a * m))
    ll = -0.5 * (d * logdet + quad) - 0.5 * d * N * LOGTWOPI
    if approx == "fitc":
        ll -= 0.5 * d * N * LOGTWOPI        # the reference counts the constant twice (CGp.cpp:963 and :1012)
    if approx == "dtcvar":
        ll -= 0.5 * d * beta * float(np.sum(kdiag - qdiag))   # CGp.cpp:955-956 with diagD of CGp.cpp:788-791

    # G = dL/dSigma = -1/2 (d Sigma^-1 - a a'), never formed: only diag(G), B G and B G B' are needed
    diagSinv = linv - np.einsum("mn,mn->n", KufL, BS)
    gd = -0.5 * (d * diagSinv - np.sum(a * a, axis=1))
    BG = -0.5 * (d * BS - Ba @ a.T)
    BGBt = -0.5 * (d * (np.linalg.inv(Kuu) - np.linalg.inv(A)) - Ba @ Ba.T)
    if approx == "fitc":
        h = gd.copy()
        gKuf = 2.0 * (BG - B * gd[None, :])
        gKuu = -BGBt + (B * gd[None, :]) @ B.T
    else:
        h = np.zeros(N)
        gKuf = 2.0 * BG
        gKuu = -BGBt
    gbeta = -float(np.sum(gd)) / (beta * beta)
    if approx == "dtcvar":
        # penalty -1/2 d beta (sum_i k_ii - tr(K_uu^-1 K_uf K_fu))
        c = -0.5 * d * beta
        h = h + c
        gKuf = gKuf - 2.0 * c * B
        gKuu = gKuu + c * (B @ B.T)
        gbeta += -0.5 * d * float(np.sum(kdiag - qdiag))
    gKuu = 0.5 * (gKuu + gKuu.T)

    # kernel parameters: K_uu (symmetric, white on its diagonal), K_uf (cross), diag(K) (FITC / DTCVAR)
    gk = O.kern_grad_params(kern, Xu, gKuu) + O.kern_grad_params(kern, Xu, gKuf, X2=X)
    if approx in ("fitc", "dtcvar"):
        gk = gk + O.kern_grad_params(kern, X, np.diag(h))
    pos = 0
    for t, p in kern:                       # natural -> transformed (CKern.cpp:50-63)
        for i, kd in enumerate(O.transform_kinds(t, D)):
            gk[pos + i] *= O.gradfact(float(p[i]), kd)
        pos += O.nparams(t, D)
    # inducing inputs
    gXu = _sym_gradX(kern, Xu, gKuu) + np.einsum("ink,in->ik", O.kern_gradX(kern, Xu, X), gKuf)
    glogbeta = gbeta * beta                 # exp transform of beta (CGp.cpp:1073-1076, CTransform.cpp:25-53)
    g = np.concatenate([gXu.T.reshape(-1), gk, [glogbeta]])
    return dict(ll=ll, g=g, gXu=gXu, gk=gk, gbeta=glogbeta)


def sparse_posterior(kern, X, y, Xu, beta, Xs, approx="dtc", bias=None, scale=None):
    """CGp::posteriorMeanVar for the sparse approximations: updateAlpha (CGp.cpp:490-521) and the sparse branch of
    _posteriorVar (CGp.cpp:584-599), then output scale and bias (CGp.cpp:561-573, 618-626).  With Lambda as above and
    A = K_uu + K_uf Lambda^-1 K_fu (beta times the reference's A):
        mean = k_*u' A^-1 K_uf Lambda^-1 m,      var = k_** - k_*u' (K_uu^-1 - A^-1) k_*u + 1/beta
    (k_** through diagComputeElement, i.e. with the white noise; the 1/beta is added for every approximation;
    DTCVAR predicts like DTC)."""
    X = np.asarray(X, dtype=np.float64)
    Xu = np.asarray(Xu, dtype=np.float64)
    Xs = np.asarray(Xs, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(X.shape[0], -1)
    N, d = y.shape
    bias = np.zeros(d) if bias is None else np.asarray(bias, dtype=np.float64).reshape(d)
    scale = np.ones(d) if scale is None else np.asarray(scale, dtype=np.float64).reshape(d)
    m = (y - bias[None, :]) / scale[None, :]
    Kuu = O.kern_compute(kern, Xu)
    Kuf = O.kern_cross(kern, Xu, X)
    lam = np.full(N, 1.0 / beta)
    if approx.lower() == "fitc":
        lam = lam + (O.kern_diag(kern, X) - np.einsum("ij,ij->j", Kuf, np.linalg.solve(Kuu, Kuf)))
    KufL = Kuf / lam[None, :]
    A = Kuu + KufL @ Kuf.T
    A = 0.5 * (A + A.T)
    alpha = np.linalg.solve(A, KufL @ m)                 # M x d
    ksu = O.kern_cross(kern, Xu, Xs)                     # M x Ns
    mu = ksu.T @ alpha
    v = O.kern_diag(kern, Xs) - np.einsum("ij,ij->j", ksu, np.linalg.solve(Kuu, ksu) - np.linalg.solve(A, ksu)) + 1.0 / beta
    return mu * scale[None, :] + bias[None, :], v[:, None] * (scale * scale)[None, :]
