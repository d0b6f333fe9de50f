"""The C++ host side of the drop-in (gpc_b200/cpp: CGpB200 : CGp, CGplvmB200 : CGplvm) without a GPU: what must hold on
any machine.  The driver oracle/_ref/cgp_b200_check (tests/cpp/cgp_b200_check.cpp, built by oracle/build_ref.sh against
the unmodified reference) runs the reference class and the drop-in class on the same data in one process."""
import json
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CHECK = os.path.join(ROOT, "oracle", "_ref", "cgp_b200_check")


def _run(*args, expect_ok=True):
    if not os.path.exists(CHECK):
        pytest.skip("oracle/_ref/cgp_b200_check not built (python __graft_entry__.py in the build container)")
    out = subprocess.run([CHECK] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    if expect_ok:
        assert out.returncode == 0, out.stderr[-2000:]
        lines = [l for l in out.stdout.splitlines() if not l.startswith("Warning:")]
        return json.loads("\n".join(lines))
    return out


@pytest.mark.parametrize("mode,N,D,d", [("gp", 60, 2, 1), ("gp", 45, 3, 2), ("gplvm", 50, 2, 4)])
def test_component_outside_the_device_path_falls_through_to_the_reference_host_code(mode, N, D, d):
    """ratquad is not a device kernel: every call of the drop-in class must land in the inherited reference
    implementation -- bit-identical results, no device evaluation, no CUDA needed."""
    r = _run(mode, N, D, d, 3, "ratquad,bias,white", 1 if (d == 1 or mode == "gplvm") else 0, 0, 12)
    assert r["on_device"] == 0 and r["device_evals"] == 0
    assert r["ll_ref"] == r["ll_dev"] == r["ll_dev_again"]
    for k in ("g", "out", "std", "opt"):
        assert r[k + "_ref"] == r[k + "_dev"], k
    assert r["opt_ll_ref"] == r["opt_ll_dev"]
    assert r["opt_ref"] != r["g_ref"] and r["opt_ll_ref"] > r["ll_ref"]  # the optimiser did move


def test_device_path_fails_loudly_without_a_gpu():
    """No CPU fallback for a model the device path covers: without a CUDA device the drop-in class throws
    ndlexceptions::Error carrying gpc_last_error() (it never silently computes on the host)."""
    import ctypes
    try:
        ndev = ctypes.CDLL(os.path.join(ROOT, "gpc_b200", "libgpc_b200.so")).gpc_device_count()
    except OSError:
        pytest.skip("libgpc_b200.so not built")
    if ndev > 0:
        pytest.skip("a CUDA device is present")
    out = _run("gp", 40, 2, 1, 3, "rbf,bias,white", 0, 0, 0, expect_ok=False)
    assert out.returncode == 1
    assert "gpc_b200:" in out.stderr


def _sinc_svml(path):
    import numpy as np
    f = np.load(os.path.join(HERE, "golden", "gp_reference.npz"))
    X, y = f["sinc_X"], np.asarray(f["sinc_y"]).ravel()
    with open(path, "w") as fh:
        for i in range(len(y)):
            fh.write("%.17g 1:%.17g\n" % (y[i], X[i, 0]))


def test_unmodified_front_end_on_the_drop_in_class_learn_and_gnuplot(tmp_path):
    """The reference's gp.cpp compiled with `-include gp_dropin.h` (oracle/_ref/gp_l2) next to the plain build: `gp learn`
    with a kernel outside the device path writes the same model file, and `gp gnuplot` on the model it reads back
    (readGpB200FromFile -> CGpB200() -> fromStream, then out() through the private noise pointer) writes the same
    plot data -- the whole front-end flow, constructors, stream I/O and the virtual dispatch, on any machine."""
    gp, gp_l2 = (os.path.join(ROOT, "oracle", "_ref", n) for n in ("gp", "gp_l2"))
    if not (os.path.exists(gp) and os.path.exists(gp_l2)):
        pytest.skip("oracle/_ref/gp, gp_l2 not built")
    _sinc_svml(str(tmp_path / "sinc.svml"))
    for exe, tag in ((gp, "ref"), (gp_l2, "l2")):
        out = subprocess.run([exe, "-v", "1", "-s", "7", "learn", "-k", "ratquad", "-#", "30", "sinc.svml", "m_" + tag],
                             cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
        out = subprocess.run([exe, "gnuplot", "sinc.svml", "m_" + tag, "plot_" + tag], cwd=str(tmp_path),
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    body = lambda p: open(str(tmp_path / p)).read().split("\n", 1)[1]   # line 1 records the command line
    assert body("m_ref") == body("m_l2")
    for part in ("line_data", "error_bar_data", "scatter_data"):
        assert open(str(tmp_path / ("plot_ref_%s.dat" % part))).read() == \
            open(str(tmp_path / ("plot_l2_%s.dat" % part))).read(), part


@pytest.mark.parametrize("spec,D,prior", [("rbf,bias,white", 1, 0), ("rbfard,matern52,lin,poly,bias,white", 3, 1),
                                          ("matern32,rbf,white", 2, 1)])
def test_kernel_bridge_flattening_and_gradient_finish(spec, D, prior):
    """GpcKernBridge on the host: component types / natural parameters in the reference's order, and finishGradient()
    (each component's prior gradient, then the transform factor) reproduces CKern::getGradTransParams (CKern.cpp:50-63)
    from the natural, prior-free gradient -- which is what the device returns."""
    import numpy as np
    r = _run("bridge", 30, D, 0, 5, spec, 0, prior)
    code = {"white": 0, "bias": 1, "rbf": 2, "rbfard": 3, "matern32": 4, "matern52": 5, "lin": 6, "poly": 7}
    assert r["supported"] == 1
    assert r["types"] == [code[k] for k in spec.split(",")]
    assert r["params"] == r["kern_params"] and len(r["params"]) == r["nparams"]
    a, b = np.array(r["g_bridge"]), np.array(r["g_ref"])
    assert np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))) < 1e-14


def test_unmodified_gplvm_front_end_on_the_drop_in_class_falls_through(tmp_path):
    """gplvm.cpp compiled with `-include gplvm_dropin.h` (oracle/_ref/gplvm_l2): with the mlp kernel (not a device
    kernel) `gplvm learn` must write the same model as the plain build -- constructors, PCA initialisation, the
    optimiser's virtual dispatch and the stream writer all go through CGplvmB200."""
    import numpy as np
    a, b = (os.path.join(ROOT, "oracle", "_ref", n) for n in ("gplvm", "gplvm_l2"))
    if not (os.path.exists(a) and os.path.exists(b)):
        pytest.skip("oracle/_ref/gplvm, gplvm_l2 not built")
    Y = np.load(os.path.join(HERE, "golden", "oil_train.npz"))["Y"][:100]
    with open(str(tmp_path / "oil100.svml"), "w") as f:
        for i in range(Y.shape[0]):
            f.write("0 " + " ".join("%d:%.17g" % (j + 1, Y[i, j]) for j in range(Y.shape[1])) + "\n")
    for exe, tag in ((a, "ref"), (b, "l2")):
        out = subprocess.run([exe, "-v", "1", "-s", "1", "learn", "-k", "mlp", "-#", "15", "oil100.svml", "m_" + tag],
                             cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    body = lambda p: open(str(tmp_path / p)).read().split("\n", 1)[1]
    assert body("m_ref") == body("m_l2")


def test_makefile_builds_the_host_classes_against_the_reference_tree(tmp_path):
    """gpc_b200/cpp/Makefile: what INTEGRATION.md tells a maintainer to run.  Needs the reference sources, so it runs in
    the build container only."""
    ref = os.environ.get("GPC_REFERENCE", "/root/reference")
    if not os.path.isdir(ref):
        pytest.skip("reference sources not present")
    blasmap = os.path.join(ROOT, "oracle", "_ref", "obj", "blasmap.h")
    extra = ("-include " + blasmap) if os.path.exists(blasmap) else ""
    out = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "gpc_b200", "cpp"), "BUILD=" + str(tmp_path),
                          "GPC_REFERENCE=" + ref, "EXTRA_CXXFLAGS=" + extra, "all", "frontends"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    for f in ("libgpc_cgp.a", "gp_dropin.o", "gplvm_dropin.o"):
        assert os.path.getsize(str(tmp_path / f)) > 0
    syms = subprocess.run(["nm", "-C", str(tmp_path / "gp_dropin.o")], capture_output=True, text=True).stdout
    assert "CGpB200::CGpB200(CKern*, CNoise*, CMatrix*, int, unsigned int, int)" in syms    # gp.cpp:392 now builds the drop-in
    assert "readGpB200FromFile" in syms


# ---- the DEVICE-path logic of the C++ host classes, driven through a host test double of the library -------------------
MOCK = os.path.join(ROOT, "oracle", "_ref", "mock")


def _run_mock(*args):
    """cgp_b200_check with tests/cpp/mock_gpc_b200.cpp (the C ABI implemented with the reference's own classes) placed in
    front of libgpc_b200.so: uploads, caching, gradient assembly and error mapping of CGpB200 / CGplvmB200 run without a
    GPU.  What it cannot show is the device arithmetic -- that is tests/test_gpu_cpp_host.py."""
    if not (os.path.exists(CHECK) and os.path.exists(os.path.join(MOCK, "libgpc_b200.so"))):
        pytest.skip("oracle/_ref/cgp_b200_check or the mock library not built")
    env = dict(os.environ, LD_LIBRARY_PATH=MOCK + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    out = subprocess.run([CHECK] + [str(a) for a in args], capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "mock_gpc_b200: host test double in use" in out.stderr
    lines = [l for l in out.stdout.splitlines() if not l.startswith("Warning:")]
    return json.loads("\n".join(lines))


def _rel(a, b):
    import numpy as np
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), 1.0)))


@pytest.mark.parametrize("mode,N,D,d,kern,scale,prior,iters", [
    ("gp", 40, 1, 1, "rbf,bias,white", 0, 0, 15),
    ("gp", 90, 3, 1, "rbf,lin,bias,white", 0, 0, 12),
    ("gp", 80, 4, 2, "rbfard,bias,white", 0, 0, 12),
    ("gp", 70, 2, 1, "rbf,white", 1, 0, 12),            # learnt output scale: m keeps its old scale (CGp.cpp:429-437)
    ("gp", 60, 3, 1, "matern52,poly,white", 0, 1, 12),  # a gamma prior on the first parameter
    ("gplvm", 50, 2, 5, "rbf,bias,white", 0, 0, 8),
    ("gplvm", 45, 3, 4, "rbfard,white", 1, 0, 0),
    ("gplvm", 40, 2, 6, "matern52,lin,white", 1, 1, 6),
])
def test_device_path_logic_of_the_host_classes_through_the_test_double(mode, N, D, d, kern, scale, prior, iters):
    r = _run_mock(mode, N, D, d, 7, kern, scale, prior, iters)
    assert r["on_device"] == 1
    assert r["evals_first"] == 1          # gradient + two likelihood calls at one point: ONE evaluation
    assert r["ll_dev"] == r["ll_dev_again"]
    assert _rel(r["ll_ref"], r["ll_dev"]) <= 1e-10
    assert _rel(r["g_ref"], r["g_dev"]) <= 1e-8
    assert _rel(r["out_ref"], r["out_dev"]) <= 1e-9 and _rel(r["std_ref"], r["std_dev"]) <= 1e-9
    if iters:
        assert r["opt_ll_ref"] > r["ll_ref"] and r["device_evals"] > iters
        assert _rel(r["opt_ref"], r["opt_dev"]) <= 1e-5 and _rel(r["opt_ll_ref"], r["opt_ll_dev"]) <= 1e-6


def test_download_accessors_of_cgp_b200_through_the_test_double():
    """CGpB200::downloadK / downloadInvK / downloadLcholK / downloadAlpha (the -DDBG view of CGp.h:359-361): refused before
    anything was evaluated, right shapes and symmetry flags, and the identities K K^-1 = I, L L' = K, alpha = K^-1 m."""
    if not (os.path.exists(CHECK) and os.path.exists(os.path.join(MOCK, "libgpc_b200.so"))):
        pytest.skip("oracle/_ref/cgp_b200_check or the mock library not built")
    env = dict(os.environ, LD_LIBRARY_PATH=MOCK + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    out = subprocess.run([CHECK, "download", "50", "3", "2", "5", "rbf,lin,bias,white"], capture_output=True, text=True,
                         timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-1500:]
    r = json.loads(out.stdout)
    assert r["refused_before_eval"] == 1
    assert r["rows"] == [50, 50, 50, 50] and r["cols"] == [50, 50, 50, 2] and r["symmetric_flags"] == [1, 1, 0]
    assert r["err_K"] <= 1e-12 and r["err_KinvK"] <= 1e-10 and r["err_LLt"] <= 1e-12 and r["err_alpha"] <= 1e-10
    assert r["err_upper"] == 0.0


def test_ccmpndkern_b200_logic_through_the_test_double():
    """CCmpndKernB200 (the CKern::compute seam): bridge flattening, context reuse, cloning (the reference's own compound
    copy constructor never terminates, CKern.cpp:142-148 -- the class copies component by component) and the symmetric
    flag, with the reference's loops behind gpc_kern_build / gpc_kern_cross."""
    r = _run_mock("kern", 60, 3, 40, 3, "rbf,lin,bias,white")
    assert r["device_builds"] == 2 and r["K_symmetric"] == 1
    assert r["K_maxdiff"] == 0.0 and r["K2_maxdiff"] == 0.0 and r["clone_maxdiff"] == 0.0


def test_level1_objects_resolve_the_six_cmatrix_methods():
    """INTEGRATION.md level 1: gp_l1 / ivm_l1 are the reference's own objects with CMatrix_b200.o linked over the weakened
    CMatrix.o -- the six hot methods must come from CMatrix_b200.cpp (they reference the gpc_d* entry points)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "gp_l1")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/gp_l1 not built")
    syms = subprocess.run(["nm", "-C", "--undefined-only", exe], capture_output=True, text=True).stdout
    for fn in ("gpc_dpotrf", "gpc_dpotri", "gpc_dtrsm", "gpc_dsyrk", "gpc_dgemm", "gpc_dsymv"):
        assert fn in syms, fn
    weak = subprocess.run(["nm", "-C", os.path.join(ROOT, "oracle", "_ref", "obj_l1", "CMatrix_weak.o")], capture_output=True,
                          text=True).stdout
    assert " W CMatrix::potrf(char const*)" in weak and " W CMatrix::trsm(" in weak
