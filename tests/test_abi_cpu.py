"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/gpc_b200.h
declares, its host-only helpers (parameter counts, transforms) agree with the oracle, and compute entry points
fail loudly -- never fall back -- when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gpc_b200 as G
from gpc_b200 import _lib
from oracle import gp_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "gpc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gpc_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = G.lib()
    names = header_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), "libgpc_b200.so does not export " + n
    assert sorted(_lib.SYMBOLS) == names


def test_no_torch_types_and_citations_in_header():
    src = open(os.path.join(ROOT, "include", "gpc_b200.h")).read()
    assert "torch" not in src and "at::" not in src
    for cite in ["lapack.h:59-65", "CGp.cpp:693-712", "CMatrix.cpp:767-804", "CGp.cpp:642-663", "CKern.cpp:50-63"]:
        assert cite in src


def test_nparams_and_transforms_match_oracle():
    L = G.lib()
    for name, code in O.KERN_TYPES.items():
        for D in (1, 4, 16):
            assert L.gpc_kern_nparams(code, D) == O.nparams(name, D)
            kinds = O.transform_kinds(name, D)
            for i, kd in enumerate(kinds):
                assert L.gpc_kern_transform(code, i) == {"exp": 1, "sigmoid": 2}[kd]
    for a in [-50.0, -36.0, -3.3, 0.0, 0.7, 12.0, 36.0, 40.0]:
        for code, kd in ((1, "exp"), (2, "sigmoid")):
            x = L.gpc_transform_atox(code, a)
            assert x == pytest.approx(O.atox(a, kd), rel=1e-15)
            assert L.gpc_transform_gradfact(code, x) == pytest.approx(O.gradfact(x, kd), rel=1e-15)
    for x in [1e-6, 0.3, 0.999]:
        assert L.gpc_transform_xtoa(1, x) == pytest.approx(O.xtoa(x, "exp"), rel=1e-15)
        assert L.gpc_transform_xtoa(2, x) == pytest.approx(O.xtoa(x, "sigmoid"), rel=1e-15)


def test_host_mirror_parameter_handling(kern_mat):
    """setTransParams / getTransParams / compound ordering (CTransform.h:254-300, CKern.h:382-418)."""
    kern = G.make_kern(["rbf", "rbfard", "poly", "white"], 3)
    assert kern.getNumParams() == 2 + 5 + 3 + 1
    tp = np.linspace(-1.0, 1.0, kern.getNumParams())
    kern.setTransParams(tp)
    np.testing.assert_allclose(kern.getTransParams(), tp, rtol=0, atol=1e-14)
    ok = O.kern_from_trans(["rbf", "rbfard", "poly", "white"], tp, 3)
    np.testing.assert_allclose(kern.params, np.concatenate([p for _, p in ok]), rtol=1e-15)
    assert G.CWhiteKern(2).getParam(0) == pytest.approx(np.exp(-2.0))
    assert list(G.CRbfardKern(3).params) == [1.0, 1.0, 0.5, 0.5, 0.5]


@pytest.mark.skipif(G.lib().gpc_device_count() > 0, reason="only meaningful without a GPU")
def test_compute_fails_loudly_without_gpu():
    h = C.c_void_p()
    rc = G.lib().gpc_ctx_create(C.byref(h), 0, 128, 2, 1)
    assert rc < 0
    assert len(G.lib().gpc_last_error()) > 0
    with pytest.raises(G.GpcError):
        G.make_kern(["rbf", "white"], 2).compute(np.zeros((4, 2)))
    with pytest.raises(G.GpcError):
        G.matrix.potrf(np.eye(3))
    with pytest.raises(G.GpcError):
        G.CGp(G.make_kern(["rbf", "white"], 2), np.zeros((4, 2)), np.zeros((4, 1)))


@pytest.mark.skipif(G.lib().gpc_device_count() > 0, reason="only meaningful without a GPU")
def test_sharded_and_sparse_entry_points_fail_loudly_without_gpu():
    L = G.lib()
    h = C.c_void_p()
    devs = (C.c_int * 2)(0, 0)
    assert L.gpc_dist_create_local(C.byref(h), devs, 2, 1, 2, 1000, 3, 1, 128) < 0 and not h.value
    assert L.gpc_dist_create_nccl(C.byref(h), 0, 0, 1, None, 1, 1, 1000, 3, 1, 128) < 0 and not h.value
    assert L.gpc_sparse_create(C.byref(h), 0, 1, 1000, 50, 3, 1) < 0 and not h.value
    assert len(L.gpc_last_error()) > 0


def test_sharded_entry_points_reject_bad_shapes():
    """argument checks that need no device: P * Q must equal the number of ranks, nb a multiple of 128"""
    L = G.lib()
    h = C.c_void_p()
    devs = (C.c_int * 2)(0, 0)
    assert L.gpc_dist_create_local(C.byref(h), devs, 2, 2, 2, 1000, 3, 1, 128) == -1      # 2 x 2 grid on 2 ranks
    assert L.gpc_dist_create_local(C.byref(h), devs, 2, 1, 2, 1000, 3, 1, 100) == -1      # nb not a multiple of 128
    assert L.gpc_sparse_create(C.byref(h), 0, 3, 1000, 50, 3, 1) == -1                    # PITC (3) is not offered
    out = (C.c_int * 12)()
    assert L.gpc_dist_plan(2, 4, 8, 1000, 128, 0, out, None) == -1                         # rank out of range
    assert L.gpc_dist_plan(2, 4, 7, 1000, 128, 0, out, None) == 0 and out[0] == 8          # 8 block rows


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under gpc_b200/ may import or load it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gpc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no numpy fallback", ""), f + " references the oracle"
                assert "libgpcref" not in src


def test_fortran_shim_exports_the_lapack_h_symbols():
    """gpc_b200/libgpc_lapack_shim.so carries the five hot Fortran symbols of the reference's lapack.h:59-73, 186-222
    under their own names, so it can be linked or preloaded in front of a BLAS (level 0)."""
    import ctypes as C
    import gpc_b200
    shim = os.path.join(os.path.dirname(gpc_b200.LIB_PATH), "libgpc_lapack_shim.so")
    assert os.path.exists(shim), "run python -m gpc_b200.build"
    L = C.CDLL(shim)
    for name in ("dpotrf_", "dpotri_", "dtrsm_", "dsyrk_", "dgemm_"):
        assert hasattr(L, name), name
