"""GP model files: the native reader / writer (gpc_gp_model_read / gpc_gp_model_write,
gpc_b200/csrc/modelio.cu) against the reference's own stream code (CGp.cpp:1605-1682, CMatrix.cpp:1057-1172,
CKern.cpp:15-26, 94-126, 2668-2705, 4192-4278, CNoise.cpp:275-305).  The oracle is the compiled reference:
`cgp_b200_check modelwrite` has the reference WRITE a model and print the values it holds, `modelread` has it READ a
file and print what it then holds.  Bar: byte-identical files, bit-identical values, the reference's quirks included."""
import json
import os
import subprocess

import numpy as np
import pytest

from gpc_b200 import io

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, "oracle", "_ref")
CHECK = os.path.join(REF, "cgp_b200_check")
CODE = {"white": 0, "bias": 1, "rbf": 2, "rbfard": 3, "matern32": 4, "matern52": 5, "lin": 6, "poly": 7}


def _driver(*args, ok=True):
    if not os.path.exists(CHECK):
        pytest.skip("oracle/_ref/cgp_b200_check not built (python __graft_entry__.py in the build container)")
    out = subprocess.run([CHECK] + [str(a) for a in args], capture_output=True, text=True, timeout=300)
    if not ok:
        return out
    assert out.returncode == 0, out.stderr[-1500:]
    return json.loads(out.stdout)


def _body(path):
    return open(path).read().split("\n", 1)[1]   # line 1 is the "# comment"


def _same_values(m, ref):
    """our reader's dict against the reference's own view of a model, bit for bit"""
    for k in ("num_data", "input_dim", "output_dim", "approx_type"):
        assert m[k] == ref[k], k
    assert m["num_active"] == ref["num_active"]
    assert int(m["learn_scale"]) == ref["learn_scale"] and int(m["learn_bias"]) == ref["learn_bias"]
    assert list(m["scale"]) == ref["scale"] and list(m["bias"]) == ref["bias"]
    assert [CODE[t] for t in m["types"]] == ref["types"]
    assert [len(p) for p in m["params"]] == ref["nparams"]
    assert list(np.concatenate(m["params"])) == ref["kern_params"]
    assert m["degree"] == ref["degree"]
    assert m["noise_type"] == ref["noise_type"] and list(m["noise_params"]) == ref["noise_params"]


CASES = [("rbf,bias,white", 1, 1, 0), ("rbfard,poly,matern52,bias,white", 3, 3, 1), ("matern32,lin,white", 2, 2, 0),
         ("single:rbfard", 4, 1, 0), ("single:poly", 2, 1, 0)]


@pytest.mark.parametrize("spec,D,d,scale", CASES)
def test_reader_and_writer_against_the_reference(tmp_path, spec, D, d, scale):
    ref_file = str(tmp_path / "ref.model")
    truth = _driver("modelwrite", 25, D, d, 5, spec, scale, 0, ref_file)     # the values the reference wrote
    seen = _driver("modelread", 0, 0, 0, 0, ref_file)                         # the values the reference reads back
    m = io.read_gp_model(ref_file)
    _same_values(m, seen)                                                     # reader == the reference's reader
    assert m["top_is_cmpnd"] == (not spec.startswith("single:"))
    # writer: from the TRUE values the file must be byte-identical to the reference's
    true_model = dict(m)
    true_model["scale"], true_model["bias"] = np.array(truth["scale"]), np.array(truth["bias"])
    ours = str(tmp_path / "ours.model")
    io.write_gp_model(ours, true_model, "anything")
    assert _body(ours) == _body(ref_file)
    # and the reference reads our file exactly as it reads its own
    assert _driver("modelread", 0, 0, 0, 0, ours) == seen


def test_the_atoi_quirk_is_reproduced_and_reported(tmp_path):
    """bias 0.25 is written "0x1p-2" (no '.'): the reference reads it back through atoi as 0 (CMatrix.cpp:1081-1085)."""
    ref_file = str(tmp_path / "ref.model")
    truth = _driver("modelwrite", 25, 3, 3, 5, "rbf,white", 0, 0, ref_file)
    assert truth["bias"][1] == 0.25
    assert "0x1p-2" in open(ref_file).read()
    m = io.read_gp_model(ref_file)
    assert m["bias"][1] == 0.0 == _driver("modelread", 0, 0, 0, 0, ref_file)["bias"][1]
    assert io.gp_model_lost_values(m) == (0, 0.0)
    m["bias"] = np.array(truth["bias"])
    assert io.gp_model_lost_values(m) == (1, 0.25)


def test_reference_cli_reads_a_model_written_here(tmp_path):
    """`gp display` (gp.cpp:556-565) on a file from gpc_gp_model_write; parameters printed with 6 digits."""
    gp = os.path.join(REF, "gp")
    if not os.path.exists(gp):
        pytest.skip("oracle/_ref/gp not built")
    model = {"num_data": 40, "input_dim": 1, "output_dim": 1, "types": ["rbf", "bias", "white"],
             "params": [[1.7056, 0.9067], [0.135335], [0.0123]], "degree": [2.0, 2.0, 2.0], "scale": [1.0],
             "bias": [0.10665780624489113], "noise_params": [0.10665780624489113, 1e-6]}
    path = str(tmp_path / "m.model")
    io.write_gp_model(path, model, "written by gpc_b200")
    out = subprocess.run([gp, "display", path], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout[-1000:] + out.stderr[-1000:]
    for name, val in (("rbfinverseWidth", 1.7056), ("rbfvariance", 0.9067), ("biasvariance", 0.135335),
                      ("whitevariance", 0.0123)):
        assert "%s: %g" % (name, val) in out.stdout, out.stdout
    back = io.read_gp_model(path)
    assert back["types"] == model["types"] and list(np.concatenate(back["params"])) == [1.7056, 0.9067, 0.135335, 0.0123]
    kern = io.kern_from_model(back)
    assert kern.getNumParams() == 4 and list(kern.params) == [1.7056, 0.9067, 0.135335, 0.0123]


def test_files_the_reference_rejects_are_rejected(tmp_path):
    from gpc_b200._lib import GpcError
    ref_file = str(tmp_path / "ref.model")
    # (1) priors: the reference's writer emits a form its own reader refuses (CDist.cpp:4-10 vs 338-357)
    _driver("modelwrite", 25, 2, 1, 5, "rbf,white", 0, 1, ref_file)
    assert "numPriors=1" in open(ref_file).read()
    assert _driver("modelread", 0, 0, 0, 0, ref_file, ok=False).returncode != 0
    with pytest.raises(GpcError, match="priors"):
        io.read_gp_model(ref_file)
    # (2) a kernel outside the device path parses in the reference but has no device kernel here
    _driver("modelwrite", 25, 2, 1, 5, "ratquad,white", 0, 0, ref_file)
    with pytest.raises(GpcError, match="outside the device path"):
        io.read_gp_model(ref_file)
    # (3) truncated file, wrong field, missing file
    _driver("modelwrite", 25, 2, 1, 5, "rbf,white", 0, 0, ref_file)
    text = open(ref_file).read()
    cut = str(tmp_path / "cut.model")
    open(cut, "w").write(text[:len(text) // 2])
    with pytest.raises(GpcError):
        io.read_gp_model(cut)
    assert _driver("modelread", 0, 0, 0, 0, cut, ok=False).returncode != 0
    open(cut, "w").write(text.replace("numData=", "numdata="))
    with pytest.raises(GpcError, match="numData"):
        io.read_gp_model(cut)
    with pytest.raises(GpcError, match="cannot open"):
        io.read_gp_model(str(tmp_path / "nope.model"))


def test_comment_lines_and_carriage_returns_are_read_like_the_reference(tmp_path):
    ref_file = str(tmp_path / "ref.model")
    _driver("modelwrite", 25, 2, 2, 5, "rbf,bias,white", 1, 0, ref_file)
    text = open(ref_file).read()
    lines = text.split("\n")
    # '#' lines anywhere are skipped by ndlstrutil::getline; a '\r' at the end of a MATRIX row is dropped (CMatrix.cpp:1071)
    noisy = []
    for l in lines:
        noisy.append(l + "\r" if l.startswith(("0x", "-0x")) else l)
        if l.startswith("numCols="):
            noisy.append("# a comment in the middle")
    path = str(tmp_path / "noisy.model")
    open(path, "w").write("\n".join(noisy))
    seen = _driver("modelread", 0, 0, 0, 0, path)
    _same_values(io.read_gp_model(path), seen)
    assert seen == _driver("modelread", 0, 0, 0, 0, ref_file)


# ---- GP-LVM model files (CGplvm.cpp:761-921) ---------------------------------------------------------------------
def _same_lvm(m, ref):
    N, d, q = ref["num_data"], ref["output_dim"], ref["latent_dim"]
    assert (m["num_data"], m["output_dim"], m["latent_dim"]) == (N, d, q)
    assert int(m["latent_regularised"]) == ref["latent_regularised"]
    assert int(m["back_constrained"]) == ref["back_constrained"] == 0
    assert int(m["dynamics_learnt"]) == ref["dynamics_learnt"] == 0
    assert [CODE[t] for t in m["types"]] == ref["types"]
    assert list(np.concatenate(m["params"])) == ref["kern_params"]
    assert m["noise_type"] == ref["noise_type"] == "scale" and list(m["noise_params"]) == ref["noise_params"]
    assert list(m["Y"].reshape(-1, order="F")) == ref["Y"]       # bit for bit, column-major
    assert list(m["X"].reshape(-1, order="F")) == ref["X"]
    assert (list(m["labels"]) if m["labels"] is not None else []) == ref["labels"]


# labels always: CGplvm's labelsPresent flag is not initialised by its constructors (CGplvm.cpp:18-36), so the reference's
# behaviour without setLabels is undefined; gplvm.cpp always sets them (gplvm.cpp:583-586)
@pytest.mark.parametrize("spec,N,q,d,labels", [("rbf,bias,white", 12, 2, 3, 1), ("rbfard,white", 40, 3, 5, 1),
                                               ("matern52,lin,bias,white", 25, 2, 12, 1)])
def test_gplvm_model_files_against_the_reference(tmp_path, spec, N, q, d, labels):
    ref_file = str(tmp_path / "ref.model")
    truth = _driver("lvmwrite", N, q, d, 5, spec, labels, 0, ref_file)   # the reference builds (PCA init) and writes
    seen = _driver("lvmread", 0, 0, 0, 0, ref_file)                      # ... and reads it back
    assert truth["Y"] == seen["Y"] and truth["X"] == seen["X"]           # hexfloat + atof: this format is lossless
    m = io.read_gplvm_model(ref_file)
    _same_lvm(m, seen)
    ours = str(tmp_path / "ours.model")
    io.write_gplvm_model(ours, m, "anything")
    assert _body(ours) == _body(ref_file)                                # byte-identical to writeGplvmToFile
    assert _driver("lvmread", 0, 0, 0, 0, ours) == seen
    # the label-free form of the same file (what the writer code emits when labelsPresent is false): own round trip
    m["labels"] = None
    io.write_gplvm_model(ours, m)
    assert "labels" not in open(ours).read()
    back = io.read_gplvm_model(ours)
    assert back["labels"] is None and np.array_equal(back["Y"], m["Y"]) and np.array_equal(back["X"], m["X"])


def test_gplvm_model_files_the_library_does_not_take(tmp_path):
    from gpc_b200._lib import GpcError
    ref_file = str(tmp_path / "ref.model")
    _driver("lvmwrite", 10, 2, 3, 5, "rbf,white", 1, 0, ref_file)
    text = open(ref_file).read()
    bad = str(tmp_path / "bad.model")
    open(bad, "w").write(text.replace("dynamicsLearnt=0", "dynamicsLearnt=1"))
    with pytest.raises(GpcError, match="dynamics"):
        io.read_gplvm_model(bad)
    open(bad, "w").write(text.replace("Y:3,X:2", "Y:4,X:2"))              # FileFormatError in the reference (CGplvm.cpp:853)
    with pytest.raises(GpcError, match="Y:,X:"):
        io.read_gplvm_model(bad)
    assert _driver("lvmread", 0, 0, 0, 0, bad, ok=False).returncode != 0
    open(bad, "w").write("\n".join(text.split("\n")[:-4]))                 # rows missing
    with pytest.raises(GpcError, match="data rows"):
        io.read_gplvm_model(bad)
    with pytest.raises(GpcError):                                          # a gp file is not a gplvm file
        _driver("modelwrite", 25, 2, 1, 5, "rbf,white", 0, 0, ref_file)
        io.read_gplvm_model(ref_file)


def test_model_from_gp_writes_what_the_reference_reads(tmp_path):
    """gpc_b200.CGp.writeModelFile = io.model_from_gp + write_gp_model.  The CGp object itself needs a GPU; its host-side
    attributes are all that model_from_gp touches, so a stand-in with the same attributes is enough here.  The reference
    must read the kernel, scale, bias and Gaussian noise (bias = column means of y, variance 1e-6) back exactly."""
    import types as _types
    import gpc_b200 as G
    rng = np.random.default_rng(4)
    X, y = rng.standard_normal((30, 3)), rng.standard_normal((30, 2))
    kern = G.make_kern(["rbfard", "poly", "bias", "white"], 3, [0.1, -0.2, 0.3, -0.4, 0.5, 0.2, -0.1, 0.05, -1.0, -2.0])
    kern._components()[1].setDegree(3.0)
    gp = _types.SimpleNamespace(pkern=kern, X=X, y=y, scale=np.array([1.5, 0.7]), bias=np.array([0.3, -1.25]),
                                getNumData=lambda: 30, getOutputDim=lambda: 2)
    path = str(tmp_path / "m.model")
    io.write_gp_model(path, io.model_from_gp(gp), "from the Python mirror")
    seen = _driver("modelread", 0, 0, 0, 0, path)
    assert seen["num_data"] == 30 and seen["input_dim"] == 3 and seen["output_dim"] == 2
    assert seen["scale"] == [1.5, 0.7] and seen["bias"] == [0.3, -1.25]
    assert seen["types"] == [CODE[t] for t in ("rbfard", "poly", "bias", "white")]
    assert seen["kern_params"] == list(kern.params) and seen["degree"][1] == 3.0
    assert seen["noise_type"] == "gaussian" and seen["noise_params"] == list(y.mean(axis=0)) + [1e-6]
    back = io.read_gp_model(path)
    k2 = io.kern_from_model(back)
    assert list(k2.params) == list(kern.params) and k2._components()[1].getDegree() == 3.0
