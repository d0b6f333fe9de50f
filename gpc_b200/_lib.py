"""ctypes binding of libgpc_b200.so (include/gpc_b200.h).  No fallbacks: if the CUDA library is missing or
there is no CUDA device every compute call raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpc_b200.so")
_lib = None

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
i64 = C.c_int64


class KComp(C.Structure):
    _fields_ = [("type", C.c_int), ("nparams", C.c_int), ("params", c_double_p), ("degree", C.c_double)]


GPC_MAX_COMPONENTS, GPC_MAX_PARAMS, GPC_MODEL_MAX_OUT = 16, 288, 256


class KernSpec(C.Structure):
    """gpc_kern_spec of include/gpc_b200.h"""
    _fields_ = [("top_is_cmpnd", C.c_int), ("input_dim", C.c_int), ("ncomp", C.c_int),
                ("type", C.c_int * GPC_MAX_COMPONENTS), ("nparams", C.c_int * GPC_MAX_COMPONENTS),
                ("degree", C.c_double * GPC_MAX_COMPONENTS), ("params", C.c_double * GPC_MAX_PARAMS)]


class NoiseSpec(C.Structure):
    """gpc_noise_spec of include/gpc_b200.h"""
    _fields_ = [("type", C.c_char * 16), ("output_dim", C.c_int), ("nparams", C.c_int),
                ("params", C.c_double * (2 * GPC_MODEL_MAX_OUT + 8))]


class GpModel(C.Structure):
    """gpc_gp_model of include/gpc_b200.h (a GP model file, CGp.cpp:1605-1682)"""
    _fields_ = [("num_data", C.c_int64), ("input_dim", C.c_int), ("output_dim", C.c_int), ("approx_type", C.c_int),
                ("num_active", C.c_uint), ("learn_scale", C.c_int), ("learn_bias", C.c_int), ("kern", KernSpec),
                ("scale", C.c_double * GPC_MODEL_MAX_OUT), ("bias", C.c_double * GPC_MODEL_MAX_OUT), ("noise", NoiseSpec)]


class GplvmModel(C.Structure):
    """gpc_gplvm_model of include/gpc_b200.h (a GP-LVM model file, CGplvm.cpp:761-898)"""
    _fields_ = [("num_data", C.c_int64), ("output_dim", C.c_int), ("latent_dim", C.c_int),
                ("latent_regularised", C.c_int), ("back_constrained", C.c_int), ("dynamics_learnt", C.c_int),
                ("has_labels", C.c_int), ("kern", KernSpec), ("noise", NoiseSpec)]


class GpcError(RuntimeError):
    pass


class MatrixNonPosDef(GpcError):
    """mirrors ndlexceptions::MatrixNonPosDef (thrown by CMatrix::potrf, CMatrix.cpp:378)"""

    def __init__(self, info):
        super().__init__("matrix is non positive definite (info=%d)" % info)
        self.info = info


OBJECTIVE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double),
                           C.POINTER(C.c_double))

# every symbol include/gpc_b200.h declares (tests/test_abi_cpu.py checks the .so exports all of them)
SYMBOLS = [
    "gpc_last_error", "gpc_device_count", "gpc_kern_nparams", "gpc_kern_transform", "gpc_transform_atox",
    "gpc_transform_xtoa", "gpc_transform_gradfact", "gpc_ctx_create", "gpc_ctx_destroy", "gpc_ctx_set_stream",
    "gpc_ctx_get_stream", "gpc_ctx_sync", "gpc_ctx_launch_count", "gpc_set_X", "gpc_set_M", "gpc_set_Y",
    "gpc_kern_build", "gpc_kern_cross", "gpc_kern_diag", "gpc_add_diag", "gpc_potrf", "gpc_jitchol",
    "gpc_solve_alpha", "gpc_inverse", "gpc_alpha_from_inverse", "gpc_grad", "gpc_kern_grad", "gpc_kern_grad_cross", "gpc_posterior",
    "gpc_eval", "gpc_download", "gpc_last_timings", "gpc_last_enqueue_ms", "gpc_dpotrf", "gpc_dpotri", "gpc_dtrsm", "gpc_dsyrk",
    "gpc_dgemm", "gpc_dsymv", "gpc_dsyr", "gpc_bench_dmma_peak", "gpc_bench_imma_peak", "gpc_bench_imma_peak_sustained", "gpc_bench_syrk", "gpc_ctx_set_profile", "gpc_last_gemm_profile", "gpc_last_gemm_profile_split", "gpc_bench_gemm",
    "gpc_bench_leaf", "gpc_bench_oz_stamps", "gpc_ctx_dims", "gpc_gp_optimise_scg", "gpc_scg_minimise", "gpc_svml_dims", "gpc_svml_read", "gpc_set_gemm_engine", "gpc_gemm_engine_slices", "gpc_gemm_check", "gpc_oz_slice_check", "gpc_oz_wave_cuts",
    "gpc_gp_model_read", "gpc_gp_model_write", "gpc_gp_model_check_roundtrip", "gpc_gplvm_model_read",
    "gpc_gplvm_model_write",
    "gpc_sparse_create", "gpc_sparse_destroy", "gpc_sparse_set_data", "gpc_sparse_eval", "gpc_sparse_posterior",
    "gpc_sparse_launch_count",
    "gpc_dist_unique_id", "gpc_dist_create_nccl", "gpc_dist_create_local", "gpc_dist_destroy", "gpc_dist_set_data",
    "gpc_dist_eval", "gpc_dist_download_kinv", "gpc_dist_info", "gpc_dist_plan",
]


def lib():
    """Load the CUDA library; raises if it has not been built (python -m gpc_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GpcError("libgpc_b200.so not built: run `python -m gpc_b200.build` (needs nvcc); there is no CPU path")
    L = C.CDLL(LIB_PATH)
    L.gpc_last_error.restype = C.c_char_p
    L.gpc_transform_atox.restype = C.c_double
    L.gpc_transform_atox.argtypes = [C.c_int, C.c_double]
    L.gpc_transform_xtoa.restype = C.c_double
    L.gpc_transform_xtoa.argtypes = [C.c_int, C.c_double]
    L.gpc_transform_gradfact.restype = C.c_double
    L.gpc_transform_gradfact.argtypes = [C.c_int, C.c_double]
    L.gpc_ctx_get_stream.restype = C.c_void_p
    L.gpc_ctx_get_stream.argtypes = [C.c_void_p]
    L.gpc_ctx_launch_count.restype = i64
    L.gpc_ctx_launch_count.argtypes = [C.c_void_p]
    L.gpc_ctx_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, i64, C.c_int, C.c_int]
    L.gpc_ctx_destroy.argtypes = [C.c_void_p]
    L.gpc_ctx_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.gpc_ctx_sync.argtypes = [C.c_void_p]
    L.gpc_set_X.argtypes = [C.c_void_p, C.c_void_p, i64, C.c_int, i64]
    L.gpc_set_M.argtypes = [C.c_void_p, C.c_void_p, i64, C.c_int, i64]
    L.gpc_set_Y.argtypes = [C.c_void_p, C.c_void_p, i64, C.c_int, i64, C.c_void_p, C.c_void_p]
    L.gpc_kern_build.argtypes = [C.c_void_p, C.POINTER(KComp), C.c_int]
    L.gpc_kern_cross.argtypes = [C.c_void_p, C.POINTER(KComp), C.c_int, C.c_void_p, i64, i64, C.c_void_p, i64]
    L.gpc_kern_diag.argtypes = [C.c_void_p, C.POINTER(KComp), C.c_int, C.c_void_p, i64, i64, C.c_void_p]
    L.gpc_add_diag.argtypes = [C.c_void_p, C.c_double]
    L.gpc_potrf.argtypes = [C.c_void_p, c_int_p, c_double_p]
    L.gpc_jitchol.argtypes = [C.c_void_p, C.c_int, c_double_p, c_double_p]
    L.gpc_solve_alpha.argtypes = [C.c_void_p, c_double_p]
    L.gpc_inverse.argtypes = [C.c_void_p]
    L.gpc_alpha_from_inverse.argtypes = [C.c_void_p, c_double_p]
    L.gpc_grad.argtypes = [C.c_void_p, C.POINTER(KComp), C.c_int, C.c_void_p, C.c_void_p]
    L.gpc_kern_grad.argtypes = [C.c_void_p, C.POINTER(KComp), C.c_int, C.c_void_p, i64, C.c_void_p, C.c_void_p]
    L.gpc_kern_grad_cross.argtypes = [C.c_void_p, C.POINTER(KComp), C.c_int, C.c_void_p, i64, i64, C.c_void_p, i64, C.c_void_p,
                                      C.c_void_p]
    L.gpc_posterior.argtypes = [C.c_void_p, C.POINTER(KComp), C.c_int, C.c_void_p, i64, i64, C.c_void_p, C.c_void_p]
    L.gpc_eval.argtypes = [C.c_void_p, C.POINTER(KComp), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gpc_download.argtypes = [C.c_void_p, C.c_int, C.c_void_p, i64]
    L.gpc_last_timings.argtypes = [C.c_void_p, C.c_void_p]
    L.gpc_last_enqueue_ms.argtypes = [C.c_void_p, c_double_p]
    L.gpc_dpotrf.argtypes = [C.c_int, C.c_char, i64, C.c_void_p, i64, c_int_p]
    L.gpc_dpotri.argtypes = [C.c_int, C.c_char, i64, C.c_void_p, i64, c_int_p]
    L.gpc_dtrsm.argtypes = [C.c_int, C.c_char, C.c_char, C.c_char, C.c_char, i64, i64, C.c_double, C.c_void_p, i64,
                            C.c_void_p, i64]
    L.gpc_dsyrk.argtypes = [C.c_int, C.c_char, C.c_char, i64, i64, C.c_double, C.c_void_p, i64, C.c_double,
                            C.c_void_p, i64]
    L.gpc_dgemm.argtypes = [C.c_int, C.c_char, C.c_char, i64, i64, i64, C.c_double, C.c_void_p, i64, C.c_void_p, i64,
                            C.c_double, C.c_void_p, i64]
    L.gpc_dsyr.argtypes = [C.c_int, C.c_char, i64, C.c_double, C.c_void_p, i64, C.c_void_p, i64]
    L.gpc_dsymv.argtypes = [C.c_int, C.c_char, i64, C.c_double, C.c_void_p, i64, C.c_void_p, C.c_double, C.c_void_p]
    L.gpc_ctx_set_profile.argtypes = [C.c_void_p, C.c_int]
    L.gpc_last_gemm_profile.argtypes = [C.c_void_p, c_double_p, C.POINTER(i64), c_double_p]
    L.gpc_last_gemm_profile_split.argtypes = [C.c_void_p, C.c_void_p]
    L.gpc_bench_dmma_peak.argtypes = [C.c_int, c_double_p]
    L.gpc_bench_imma_peak.argtypes = [C.c_int, C.c_int, c_double_p]
    L.gpc_bench_imma_peak_sustained.argtypes = [C.c_int, C.c_int, C.c_double, c_double_p]
    L.gpc_bench_gemm.argtypes = [C.c_int, i64, i64, i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p]
    L.gpc_bench_syrk.argtypes = [C.c_int, i64, i64, C.c_int, c_double_p]
    L.gpc_set_gemm_engine.argtypes = [C.c_int, C.c_int, i64, i64]
    L.gpc_ctx_dims.argtypes = [C.c_void_p, C.POINTER(i64), c_int_p, c_int_p]
    L.gpc_gp_optimise_scg.argtypes = [C.c_void_p, C.POINTER(KComp), C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p,
                                      c_int_p, c_int_p]
    L.gpc_scg_minimise.argtypes = [OBJECTIVE_FN, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double,
                                   C.c_void_p, c_int_p, c_int_p]
    L.gpc_svml_dims.argtypes = [C.c_char_p, C.POINTER(i64), c_int_p]
    L.gpc_svml_read.argtypes = [C.c_char_p, C.c_void_p, i64, C.c_void_p, i64, C.c_int]
    L.gpc_gp_model_read.argtypes = [C.c_char_p, C.POINTER(GpModel)]
    L.gpc_gp_model_write.argtypes = [C.c_char_p, C.POINTER(GpModel), C.c_char_p]
    L.gpc_gp_model_check_roundtrip.argtypes = [C.POINTER(GpModel), c_int_p, c_double_p]
    L.gpc_gplvm_model_read.argtypes = [C.c_char_p, C.POINTER(GplvmModel), C.c_void_p, i64, C.c_void_p, i64, C.c_void_p]
    L.gpc_gplvm_model_write.argtypes = [C.c_char_p, C.POINTER(GplvmModel), C.c_void_p, i64, C.c_void_p, i64, C.c_void_p,
                                        C.c_char_p]
    L.gpc_bench_oz_stamps.argtypes = [C.c_int, i64, i64, i64, C.c_int, C.c_void_p]
    L.gpc_bench_leaf.argtypes = [C.c_int, C.c_int, c_double_p, C.c_void_p]
    L.gpc_oz_slice_check.argtypes = [C.c_int, i64, i64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gpc_gemm_check.argtypes = [C.c_int, i64, i64, i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                 C.c_void_p, C.c_void_p, C.c_void_p]
    L.gpc_sparse_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, i64, C.c_int, C.c_int, C.c_int]
    L.gpc_sparse_destroy.argtypes = [C.c_void_p]
    L.gpc_sparse_set_data.argtypes = [C.c_void_p, C.c_void_p, i64, C.c_void_p, i64]
    L.gpc_sparse_eval.argtypes = [C.c_void_p, C.POINTER(KComp), C.c_int, C.c_void_p, i64, C.c_double, C.c_void_p, C.c_void_p,
                                  C.c_void_p, c_double_p]
    L.gpc_sparse_posterior.argtypes = [C.c_void_p, C.POINTER(KComp), C.c_int, C.c_void_p, i64, i64, C.c_void_p, C.c_void_p]
    L.gpc_sparse_launch_count.restype = i64
    L.gpc_sparse_launch_count.argtypes = [C.c_void_p]
    L.gpc_dist_unique_id.argtypes = [C.c_void_p]
    L.gpc_dist_create_nccl.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, i64,
                                       C.c_int, C.c_int, C.c_int]
    L.gpc_dist_create_local.argtypes = [C.POINTER(C.c_void_p), c_int_p, C.c_int, C.c_int, C.c_int, i64, C.c_int, C.c_int,
                                        C.c_int]
    L.gpc_dist_destroy.argtypes = [C.c_void_p]
    L.gpc_dist_set_data.argtypes = [C.c_void_p, C.c_void_p, i64, C.c_void_p, i64]
    L.gpc_dist_eval.argtypes = [C.c_void_p, C.POINTER(KComp), C.c_int, C.c_void_p, C.c_void_p]
    L.gpc_dist_download_kinv.argtypes = [C.c_void_p, C.c_void_p, i64]
    L.gpc_dist_info.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.gpc_dist_plan.argtypes = [C.c_int, C.c_int, C.c_int, i64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    _lib = L
    return L


def check(rc):
    """<0 -> GpcError with the library's message; >0 (numerical info) is returned to the caller."""
    if rc < 0:
        raise GpcError(lib().gpc_last_error().decode() or ("gpc error %d" % rc))
    return rc


def fmat(a):
    """column-major fp64 view/copy (CMatrix layout, CMatrix.h:268)"""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    return np.asfortranarray(a)


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)
