"""tests/golden/make_golden.py -- regenerates the committed golden fixtures.  Run in the BUILD container
(needs /root/reference and oracle/_ref):   python tests/golden/make_golden.py

Two sources, both from the reference itself (nothing here is produced by gpc_b200):
  (1) the reference's own MATLAB-generated known answers in /root/reference/matfiles/*.mat
      (testKern.cpp:236-376, testGp.cpp:118-150, testMatrix.cpp:187-236, 332-393, 606-836), re-packed as npz;
  (2) outputs of the unmodified reference compiled by oracle/build_ref.sh (oracle/_ref/libgpcref.so)
      on seeded random inputs, for cases the MATLAB fixtures do not cover (compound kernels of the in-scope
      components, CGp ll / gradient / posterior at several N, CGplvm ll / gradient, jitChol).
"""
import os
import sys

import numpy as np
import scipy.io as sio

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refbind as R  # noqa: E402

MF = os.environ.get("GPC_REFERENCE", "/root/reference") + "/matfiles/"
EX = os.environ.get("GPC_REFERENCE", "/root/reference") + "/examples/"


def dense(a):
    return a.toarray() if hasattr(a, "toarray") else np.asarray(a, dtype=np.float64)


def load_svml(path):
    """SVM-light reader, same semantics as CClctrl::readSvmlDataFile (CClctrl.cpp:55-171):
    'label idx:val idx:val ...', 1-based feature indices, '#' comments."""
    rows, ys, maxf = [], [], 0
    for line in open(path):
        line = line.split("#")[0].strip()
        if not line:
            continue
        tok = line.split()
        ys.append(float(tok[0]))
        feats = {}
        for t in tok[1:]:
            i, v = t.split(":")
            feats[int(i)] = float(v)
            maxf = max(maxf, int(i))
        rows.append(feats)
    X = np.zeros((len(rows), maxf))
    for r, f in enumerate(rows):
        for i, v in f.items():
            X[r, i - 1] = v
    return X, np.array(ys)[:, None]


def kern_fixtures():
    out = {}
    for name in ["rbf", "rbfard", "matern32", "matern52", "lin", "poly", "white", "bias"]:
        m = sio.loadmat(MF + name + "KernTest.mat")
        G2 = m["G2"].ravel()
        out.update({
            name + "_X": m["X"], name + "_X2": m["X2"], name + "_params": m["params"].ravel().astype(np.float64),
            name + "_K2": dense(m["K2"]), name + "_K4": dense(m["K4"]), name + "_k2": dense(m["k2"]).ravel(),
            name + "_covGrad": dense(m["covGrad"]), name + "_covGrad2": dense(m["covGrad2"]),
            name + "_g2": m["g2"].ravel().astype(np.float64), name + "_g4": m["g4"].ravel().astype(np.float64),
            # d k(X_i, X2_:)/d X_i for the first 6 rows i (full cell is 100 x (200 x 4))
            name + "_G2": np.stack([dense(G2[i]) for i in range(6)]),
            name + "_GD2": dense(m["GD2"]).astype(np.float64),
        })
    np.savez_compressed(os.path.join(HERE, "kern_matfiles.npz"), **out)


def matrix_fixtures():
    out = {}
    for f in ["choleskyMatrixTest", "invMatrixTest", "syrkMatrixTest", "trsmMatrixTest", "gemmMatrixTest", "syrMatrixTest"]:
        m = sio.loadmat(MF + f + ".mat")
        for k, v in m.items():
            if not k.startswith("__"):
                out[f + "_" + k] = dense(v)
    np.savez_compressed(os.path.join(HERE, "matrix_matfiles.npz"), **out)


def gp_fixtures():
    m = sio.loadmat(MF + "testGpftc.mat")
    X, y = m["X"], m["y"]
    types = ["rbf", "lin", "bias", "white"]
    tp = m["params"].ravel().astype(np.float64)
    bias = y.mean(0)
    Xs = np.array([[0.1, -0.2], [1.5, 1.0], [-3.0, 3.0]])
    r = R.gp_eval(types, tp, X, y, bias=bias, Xs=Xs)
    out = dict(ftc_X=X, ftc_y=y, ftc_params=tp, ftc_bias=bias, ftc_ll_matlab=float(m["ll"]),
               ftc_grads_matlab=m["grads"].ravel(), ftc_ll_ref=r["ll"], ftc_g_ref=r["g"], ftc_Xs=Xs,
               ftc_mu_ref=r["mu"], ftc_var_ref=r["var"])
    # config 1: examples/sinc.svml with `gp learn` defaults (gp.cpp:240-349, 379-406): rbf+bias+white, theta_t=[0,0,-2,-2]
    X, y = load_svml(EX + "sinc.svml")
    tp = np.array([0.0, 0.0, -2.0, -2.0])
    Xs = np.array([[0.0], [10.0], [-2.5]])
    r = R.gp_eval(["rbf", "bias", "white"], tp, X, y, bias=y.mean(0), Xs=Xs)
    out.update(sinc_X=X, sinc_y=y, sinc_params=tp, sinc_bias=y.mean(0), sinc_ll_ref=r["ll"], sinc_g_ref=r["g"],
               sinc_Xs=Xs, sinc_mu_ref=r["mu"], sinc_var_ref=r["var"])
    np.savez_compressed(os.path.join(HERE, "gp_reference.npz"), **out)


CASES = [  # (tag, types, N, D, dout)  -- seeded random inputs run through the reference
    ("c_rbf_white", ["rbf", "white"], 300, 8, 1),
    ("c_rbfard_white", ["rbfard", "white"], 257, 5, 1),
    ("c_m52_white", ["matern52", "white"], 200, 6, 2),
    ("c_m32_bias_white", ["matern32", "bias", "white"], 129, 3, 1),
    ("c_lin_poly_white", ["lin", "poly", "white"], 150, 4, 1),
    ("c_all", ["rbf", "rbfard", "matern32", "matern52", "lin", "poly", "bias", "white"], 140, 3, 3),
]


def random_cases():
    from oracle import gp_oracle as O
    out = {}
    for idx, (tag, types, N, D, d) in enumerate(CASES):
        rng = np.random.default_rng(1000 + idx)
        X = rng.standard_normal((N, D))
        X2 = rng.standard_normal((37, D))
        y = np.sin(X[:, :1]) @ np.ones((1, d)) + 0.1 * rng.standard_normal((N, d)) + np.arange(d)[None, :]
        npar = sum(O.nparams(t, D) for t in types)
        tp = 0.5 * rng.standard_normal(npar)
        # keep noise/bias/poly terms moderate so K stays well conditioned
        pos = 0
        for t in types:
            n = O.nparams(t, D)
            if t == "white":
                tp[pos] = -2.0 + 0.3 * rng.standard_normal()
            if t in ("poly", "lin"):
                tp[pos:pos + n] = -1.5 + 0.2 * rng.standard_normal(n)
            pos += n
        bias = y.mean(0)
        scale = 1.0 + 0.5 * rng.random(d)
        cg = rng.standard_normal((N, N))
        cg = 0.5 * (cg + cg.T)
        cg2 = rng.standard_normal((N, 37))
        r = R.gp_eval(types, tp, X, y, bias=bias, scale=scale, Xs=X2)
        out.update({
            tag + "_X": X, tag + "_X2": X2, tag + "_y": y, tag + "_tparams": tp, tag + "_bias": bias,
            tag + "_scale": scale, tag + "_params": R.kern_params(types, tp, D),
            tag + "_K": R.kern_compute(types, tp, X), tag + "_Kx": R.kern_cross(types, tp, X, X2),
            tag + "_kdiag": R.kern_diag(types, tp, X2),
            tag + "_covGrad": cg, tag + "_covGrad2": cg2,
            tag + "_g": R.kern_grad(types, tp, X, cg), tag + "_g2": R.kern_grad(types, tp, X, cg2, X2),
            tag + "_gradX": R.kern_gradX(types, tp, X[:5], X2), tag + "_diagGradX": R.kern_diagGradX(types, tp, X),
            tag + "_ll": r["ll"], tag + "_gll": r["g"], tag + "_mu": r["mu"], tag + "_var": r["var"],
        })
    # CGplvm on a slice of oilTrain (config 5 shape: q=2, d=12), kernel rbf+bias+white at the gplvm defaults
    Y, _ = load_svml(EX + "oilTrain.svml")
    Y = Y[:120]
    Xl = R.gplvm_initX(Y, 2)
    tp = np.array([0.0, 0.0, -2.0, -2.0])
    r = R.gplvm_eval(["rbf", "bias", "white"], tp, Xl, Y)
    out.update(lvm_Y=Y, lvm_X=Xl, lvm_tparams=tp, lvm_ll=r["ll"], lvm_g=r["g"], lvm_m=r["m"])
    rng = np.random.default_rng(77)
    Xl2 = Xl + 0.3 * rng.standard_normal(Xl.shape)
    tp2 = np.array([0.3, -0.2, 0.1, -0.5, -1.0, -2.5])
    r = R.gplvm_eval(["rbf", "lin", "matern32", "white"], tp2, Xl2, Y)
    out.update(lvm2_X=Xl2, lvm2_tparams=tp2, lvm2_ll=r["ll"], lvm2_g=r["g"])
    # jitChol on a rank-deficient matrix (CMatrix.cpp:767-804)
    B = rng.standard_normal((40, 12))
    A = B @ B.T
    U, jit, Aj = R.jitchol(A)
    out.update(jit_A=A, jit_U=U, jit_val=jit, jit_Aout=Aj)
    np.savez_compressed(os.path.join(HERE, "random_reference.npz"), **out)


SPARSE_CASES = [  # (tag, N, D, d, kernels, approx code of CGp.h:12-19, M, beta)
    ("dtc_rbf", 60, 2, 1, "rbf,bias,white", 1, 8, 10.0),
    ("dtc_ard2", 150, 3, 2, "rbfard,lin,white", 1, 12, 25.0),
    ("dtc_poly", 120, 2, 1, "poly,rbf,bias,white", 1, 10, 5.0),
    ("fitc_rbf", 60, 2, 1, "rbf,bias,white", 2, 8, 10.0),
    ("fitc_ard2", 150, 3, 2, "rbfard,lin,white", 2, 12, 25.0),
    ("fitc_poly", 120, 2, 1, "poly,rbf,bias,white", 2, 10, 5.0),
    ("dtcvar_rbf", 60, 2, 1, "rbf,bias,white", 4, 8, 10.0),
    ("dtcvar_ard2", 150, 3, 2, "rbfard,lin,white", 4, 12, 25.0),
]


def sparse_fixtures():
    """SURVEY 8(f) row 2 (sparse approximations), for oracle/gp_sparse_oracle.py:
    (1) the reference's MATLAB known answers matfiles/testGpdtc.mat, testGpfitc.mat (testGp.cpp:21-23, 98-150): X, y,
        inducing inputs, beta, optimiser-space parameters, gradients and log-likelihood;
    (2) the compiled reference (oracle/_ref/cgp_b200_check sparse: CGp with approxType DTC / FITC / DTCVAR) on seeded
        inputs -- better conditioned than the MATLAB cases (cond(A) ~ 1e9 there) and covering two outputs, ARD, poly."""
    import json
    import subprocess
    out = {}
    for name in ("testGpdtc", "testGpfitc"):
        m = sio.loadmat(MF + name + ".mat", squeeze_me=True, struct_as_record=False)
        gi = m["gpInfoInit"]
        out.update({name + "_X": m["X"], name + "_y": np.asarray(m["y"]).reshape(-1, 1), name + "_Xu": gi.X_u,
                    name + "_beta": float(gi.beta), name + "_params": m["params"], name + "_grads": m["grads"],
                    name + "_ll": float(m["ll"]), name + "_bias": float(m["bias"]), name + "_scale": float(m["scale"]),
                    name + "_types": np.array([c.type for c in m["kernInit"].comp])})
    exe = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "cgp_b200_check")
    for tag, N, D, d, spec, approx, M, beta in SPARSE_CASES:
        r = json.loads(subprocess.run([exe, "sparse", str(N), str(D), str(d), "3", spec, str(approx), str(M), str(beta)],
                                      capture_output=True, text=True, check=True).stdout)
        out.update({tag + "_X": np.array(r["X"]).reshape(D, N).T, tag + "_y": np.array(r["y"]).reshape(d, N).T,
                    tag + "_Xu": np.array(r["X_u"]).reshape(D, M).T, tag + "_beta": r["beta"], tag + "_bias": np.array(r["bias"]),
                    tag + "_params": np.array(r["params"]), tag + "_grads": np.array(r["g"]), tag + "_ll": r["ll"],
                    tag + "_approx": np.array(r["approx"]), tag + "_types": np.array(spec.split(",")),
                    tag + "_Xs": np.array(r["Xs"]).reshape(D, -1).T, tag + "_mu": np.array(r["mu"]).reshape(d, -1).T,
                    tag + "_var": np.array(r["var"]).reshape(d, -1).T})
    np.savez_compressed(os.path.join(HERE, "sparse_reference.npz"), **out)


def ivm_fixture():
    """examples/unitsquaregp.svml (500 x 2, labels +-1): the data of the reference's IVM walk-through
    (README.md:234, `ivm learn -a 200 -k rbf examples/unitsquaregp.svml`), for tests/test_gpu_shim.py."""
    X, y = load_svml(EX + "unitsquaregp.svml")
    np.savez_compressed(os.path.join(HERE, "unitsquaregp.npz"), X=X, y=y)


if __name__ == "__main__":
    kern_fixtures()
    matrix_fixtures()
    gp_fixtures()
    random_cases()
    ivm_fixture()
    sparse_fixtures()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
