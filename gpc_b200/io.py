"""On-disk formats of the reference, read and written by the native library (no parsing is done in Python):
SVM-light input as CClctrl::readSvmlDataFile reads it (CClctrl.cpp:55-171; gpc_svml_dims / gpc_svml_read) and GP model
files as CGp::writeParamsToStream / readParamsFromStream write and read them (CGp.cpp:1605-1682; gpc_gp_model_read /
gpc_gp_model_write)."""
import ctypes as C

import numpy as np

from ._lib import GplvmModel, GpModel, check, i64, lib, ptr


def read_svml(path):
    """returns (X, y): X is N x D column-major (D = the largest feature index in the file), y is N x 1."""
    n, d = i64(0), C.c_int(0)
    check(lib().gpc_svml_dims(str(path).encode(), C.byref(n), C.byref(d)))
    X = np.zeros((n.value, d.value), order="F")
    y = np.zeros((n.value, 1), order="F")
    if n.value:
        check(lib().gpc_svml_read(str(path).encode(), ptr(X), max(n.value, 1), ptr(y), n.value, d.value))
    return X, y


def _kern_to_dict(k):
    from .kern import TYPE_NAMES
    nc = k.ncomp
    npar = [k.nparams[c] for c in range(nc)]
    flat = np.array(k.params[:sum(npar)], dtype=np.float64)
    off = np.concatenate([[0], np.cumsum(npar)]).astype(int)
    return {"top_is_cmpnd": bool(k.top_is_cmpnd), "kern_input_dim": k.input_dim,
            "types": [TYPE_NAMES[k.type[c]] for c in range(nc)],
            "params": [flat[off[c]:off[c + 1]].copy() for c in range(nc)],
            "degree": [float(k.degree[c]) for c in range(nc)]}


def _noise_to_dict(n):
    return {"noise_type": n.type.decode(), "noise_output_dim": n.output_dim,
            "noise_params": np.array(n.params[:n.nparams])}


def _kern_from_dict(k, model, default_dim):
    from .kern import TYPE_NAMES
    code = {v: kk for kk, v in TYPE_NAMES.items()}
    k.top_is_cmpnd = int(bool(model.get("top_is_cmpnd", True)))
    k.input_dim = int(model.get("kern_input_dim", default_dim))
    types = list(model["types"])
    if len(types) > len(k.type):
        raise ValueError("too many kernel components")
    k.ncomp = len(types)
    pos = 0
    for c, t in enumerate(types):
        p = np.asarray(model["params"][c], dtype=np.float64).ravel()
        k.type[c], k.nparams[c] = code[t], p.size
        k.degree[c] = float(model.get("degree", [2.0] * len(types))[c])
        if pos + p.size > len(k.params):
            raise ValueError("too many kernel parameters")
        for v in p:
            k.params[pos] = v
            pos += 1


def _noise_from_dict(n, model, default_type, d):
    n.type = str(model.get("noise_type", default_type)).encode()
    n.output_dim = int(model.get("noise_output_dim", d))
    npar = np.asarray(model["noise_params"], dtype=np.float64).ravel()
    if npar.size > len(n.params):
        raise ValueError("too many noise parameters")
    n.nparams = npar.size
    for j, v in enumerate(npar):
        n.params[j] = v


def read_gp_model(path):
    """A GP model file (what `gp learn` writes) as a dict: sizes, flags, scale, bias, kernel component types (names) with
    their NATURAL parameters and polynomial degrees, noise type and parameters.  Values are what the reference's reader
    would hold, including its atoi rule for entries without a '.' (CMatrix.cpp:1081-1085)."""
    m = GpModel()
    check(lib().gpc_gp_model_read(str(path).encode(), C.byref(m)))
    d = m.output_dim
    out = {"num_data": int(m.num_data), "input_dim": m.input_dim, "output_dim": d, "approx_type": m.approx_type,
           "num_active": int(m.num_active), "learn_scale": bool(m.learn_scale), "learn_bias": bool(m.learn_bias),
           "scale": np.array(m.scale[:d]), "bias": np.array(m.bias[:d])}
    out.update(_kern_to_dict(m.kern))
    out.update(_noise_to_dict(m.noise))
    return out


def _to_struct(model):
    m = GpModel()
    m.num_data, m.input_dim, m.output_dim = int(model["num_data"]), int(model["input_dim"]), int(model["output_dim"])
    m.approx_type, m.num_active = int(model.get("approx_type", 0)), int(model.get("num_active", 0)) & 0xFFFFFFFF
    m.learn_scale, m.learn_bias = int(bool(model.get("learn_scale", False))), int(bool(model.get("learn_bias", False)))
    _kern_from_dict(m.kern, model, model["input_dim"])
    d = m.output_dim
    if not 1 <= d <= len(m.scale):
        raise ValueError("output_dim out of range")
    for j in range(d):
        m.scale[j] = float(np.ravel(model["scale"])[j])
        m.bias[j] = float(np.ravel(model["bias"])[j])
    _noise_from_dict(m.noise, model, "gaussian", d)
    return m


def write_gp_model(path, model, comment=""):
    """Writes the dict layout of read_gp_model as a GP model file the reference's `gp display / gnuplot / relearn` read:
    byte-identical to what CGp::toFile writes for the same model (after the comment line)."""
    m = _to_struct(model)
    check(lib().gpc_gp_model_write(str(path).encode(), C.byref(m), str(comment).encode()))


def gp_model_lost_values(model):
    """(count, first) of the values the reference's reader would NOT return as written (a power of two below one is
    written "0x1p-k", has no '.', and is read back through atoi as 0): check before trusting a round trip."""
    m = _to_struct(model)
    n, first = C.c_int(0), C.c_double(0.0)
    check(lib().gpc_gp_model_check_roundtrip(C.byref(m), C.byref(n), C.byref(first)))
    return n.value, first.value


def kern_from_model(model):
    """The kernel object of a model dict: cmpnd(types...) at the file's natural parameters and polynomial degrees."""
    from .kern import make_kern
    kern = make_kern(model["types"], model["kern_input_dim"])
    for k, p, deg in zip(kern._components(), model["params"], model["degree"]):
        k.setParams(p)
        k.degree = float(deg)
    return kern


def model_from_gp(gp, noise_params=None, top_is_cmpnd=True):
    """The model dict of a gpc_b200.CGp, as `gp learn` would save it (gp.cpp:410-430): Gaussian noise whose bias is the
    column mean of y and whose variance is the CGaussianNoise default 1e-6 (CNoise.cpp), unless given."""
    comps = gp.pkern._components()
    d = gp.getOutputDim()
    if noise_params is None:
        noise_params = np.concatenate([np.asarray(gp.y).mean(axis=0), [1e-6]])
    return {
        "num_data": gp.getNumData(), "input_dim": gp.X.shape[1], "output_dim": d, "approx_type": 0, "num_active": 0,
        "learn_scale": False, "learn_bias": False, "top_is_cmpnd": top_is_cmpnd, "kern_input_dim": gp.X.shape[1],
        "types": [k.type_name for k in comps], "params": [np.array(k.params, dtype=np.float64) for k in comps],
        "degree": [float(k.degree) for k in comps], "scale": np.array(gp.scale), "bias": np.array(gp.bias),
        "noise_type": "gaussian", "noise_output_dim": d, "noise_params": np.asarray(noise_params, dtype=np.float64),
    }


def read_gplvm_model(path):
    """A GP-LVM model file (what `gplvm learn` writes, CGplvm.cpp:761-898) as a dict: the header flags, the kernel and the
    CScaleNoise parameters as in read_gp_model, the targets Y (N x d), the latent positions X (N x q) and the labels."""
    m = GplvmModel()
    check(lib().gpc_gplvm_model_read(str(path).encode(), C.byref(m), None, 0, None, 0, None))
    N, d, q = int(m.num_data), m.output_dim, m.latent_dim
    Y = np.zeros((N, d), order="F")
    X = np.zeros((N, q), order="F")
    labels = np.zeros(max(N, 1), dtype=np.int32)
    check(lib().gpc_gplvm_model_read(str(path).encode(), C.byref(m), ptr(Y), max(N, 1), ptr(X), max(N, 1), ptr(labels)))
    out = {"num_data": N, "output_dim": d, "latent_dim": q, "latent_regularised": bool(m.latent_regularised),
           "back_constrained": bool(m.back_constrained), "dynamics_learnt": bool(m.dynamics_learnt),
           "Y": Y, "X": X, "labels": labels[:N].copy() if m.has_labels else None}
    out.update(_kern_to_dict(m.kern))
    out.update(_noise_to_dict(m.noise))
    return out


def write_gplvm_model(path, model, comment=""):
    """Writes the dict layout of read_gplvm_model; byte-identical to writeGplvmToFile (CGplvm.cpp:908-921) for the same
    model (after the comment line)."""
    m = GplvmModel()
    Y = np.asfortranarray(np.asarray(model["Y"], dtype=np.float64))
    X = np.asfortranarray(np.asarray(model["X"], dtype=np.float64))
    N = Y.shape[0]
    m.num_data, m.output_dim, m.latent_dim = N, Y.shape[1], X.shape[1]
    m.latent_regularised = int(bool(model.get("latent_regularised", True)))
    m.back_constrained = int(bool(model.get("back_constrained", False)))
    m.dynamics_learnt = int(bool(model.get("dynamics_learnt", False)))
    labels = model.get("labels")
    m.has_labels = int(labels is not None)
    lab = np.ascontiguousarray(labels, dtype=np.int32) if labels is not None else None
    _kern_from_dict(m.kern, model, X.shape[1])
    _noise_from_dict(m.noise, model, "scale", Y.shape[1])
    check(lib().gpc_gplvm_model_write(str(path).encode(), C.byref(m), ptr(Y), max(N, 1), ptr(X), max(N, 1),
                                      ptr(lab) if lab is not None else None, str(comment).encode()))
