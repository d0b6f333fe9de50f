"""On-disk formats of the reference, read and written by the native library (no parsing is done in Python):
SVM-light input as CClctrl::readSvmlDataFile reads it (CClctrl.cpp:55-171; gpc_svml_dims / gpc_svml_read) and GP model
files as CGp::writeParamsToStream / readParamsFromStream write and read them (CGp.cpp:1605-1682; gpc_gp_model_read /
gpc_gp_model_write)."""
import ctypes as C

import numpy as np

from ._lib import GpModel, check, i64, lib, ptr


def read_svml(path):
    """returns (X, y): X is N x D column-major (D = the largest feature index in the file), y is N x 1."""
    n, d = i64(0), C.c_int(0)
    check(lib().gpc_svml_dims(str(path).encode(), C.byref(n), C.byref(d)))
    X = np.zeros((n.value, d.value), order="F")
    y = np.zeros((n.value, 1), order="F")
    if n.value:
        check(lib().gpc_svml_read(str(path).encode(), ptr(X), max(n.value, 1), ptr(y), n.value, d.value))
    return X, y


def read_gp_model(path):
    """A GP model file (what `gp learn` writes) as a dict: sizes, flags, scale, bias, kernel component types (names) with
    their NATURAL parameters and polynomial degrees, noise type and parameters.  Values are what the reference's reader
    would hold, including its atoi rule for entries without a '.' (CMatrix.cpp:1081-1085)."""
    from .kern import TYPE_NAMES
    m = GpModel()
    check(lib().gpc_gp_model_read(str(path).encode(), C.byref(m)))
    d, nc = m.output_dim, m.ncomp
    npar = [m.nparams[c] for c in range(nc)]
    flat = np.array(m.kern_params[:sum(npar)], dtype=np.float64)
    off = np.concatenate([[0], np.cumsum(npar)]).astype(int)
    return {
        "num_data": int(m.num_data), "input_dim": m.input_dim, "output_dim": d, "approx_type": m.approx_type,
        "num_active": int(m.num_active), "learn_scale": bool(m.learn_scale), "learn_bias": bool(m.learn_bias),
        "top_is_cmpnd": bool(m.top_is_cmpnd), "kern_input_dim": m.kern_input_dim,
        "types": [TYPE_NAMES[m.type[c]] for c in range(nc)],
        "params": [flat[off[c]:off[c + 1]].copy() for c in range(nc)],
        "degree": [float(m.degree[c]) for c in range(nc)],
        "scale": np.array(m.scale[:d]), "bias": np.array(m.bias[:d]),
        "noise_type": m.noise_type.decode(), "noise_output_dim": m.noise_output_dim,
        "noise_params": np.array(m.noise_params[:m.noise_nparams]),
    }


def _to_struct(model):
    from .kern import TYPE_NAMES
    code = {v: k for k, v in TYPE_NAMES.items()}
    m = GpModel()
    m.num_data, m.input_dim, m.output_dim = int(model["num_data"]), int(model["input_dim"]), int(model["output_dim"])
    m.approx_type, m.num_active = int(model.get("approx_type", 0)), int(model.get("num_active", 0)) & 0xFFFFFFFF
    m.learn_scale, m.learn_bias = int(bool(model.get("learn_scale", False))), int(bool(model.get("learn_bias", False)))
    m.top_is_cmpnd = int(bool(model.get("top_is_cmpnd", True)))
    m.kern_input_dim = int(model.get("kern_input_dim", model["input_dim"]))
    types = list(model["types"])
    if len(types) > len(m.type):
        raise ValueError("too many kernel components")
    m.ncomp = len(types)
    pos = 0
    for c, t in enumerate(types):
        p = np.asarray(model["params"][c], dtype=np.float64).ravel()
        m.type[c], m.nparams[c] = code[t], p.size
        m.degree[c] = float(model.get("degree", [2.0] * len(types))[c])
        if pos + p.size > len(m.kern_params):
            raise ValueError("too many kernel parameters")
        for v in p:
            m.kern_params[pos] = v
            pos += 1
    d = m.output_dim
    if not 1 <= d <= len(m.scale):
        raise ValueError("output_dim out of range")
    for j in range(d):
        m.scale[j] = float(np.ravel(model["scale"])[j])
        m.bias[j] = float(np.ravel(model["bias"])[j])
    m.noise_type = str(model.get("noise_type", "gaussian")).encode()
    m.noise_output_dim = int(model.get("noise_output_dim", d))
    npar = np.asarray(model["noise_params"], dtype=np.float64).ravel()
    if npar.size > len(m.noise_params):
        raise ValueError("too many noise parameters")
    m.noise_nparams = npar.size
    for j, v in enumerate(npar):
        m.noise_params[j] = v
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
