"""Multi-GPU exact-GP evaluation: host-side mirror of the gpc_dist_* C ABI (gpc_b200/csrc/dist.cu).

K, its Cholesky factor and K^-1 are sharded 2-D block-cyclically (nb x nb blocks over a P x Q process grid) and
K -> K^-1 happens in place in one fused right-looking sweep inside the library; this module only creates the context,
hands over the data and assembles log-likelihood / gradient like CGp.logLikelihoodGradient (CGp.cpp:913-1144).

Two back-ends of the library:
  "nccl"  : one process per GPU (torchrun).  The launcher's only job is to carry rank 0's 128-byte NCCL id to the other
            ranks: `torch.distributed` (any backend; gloo is enough) is used for exactly that broadcast.
  "local" : one process, `devices` lists the GPUs (a worker thread each inside the library, peer copies over NVLink).
            The same device may be listed several times -- the block-cyclic logic on a single GPU (tests).
There is no CPU path: without the CUDA library and a device every call raises.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, fmat, lib, ptr

GRIDS = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}   # 8 GPUs -> 2 x 4, 4 -> 2 x 2, 2 -> 1 x 2


def default_grid(world):
    if world in GRIDS:
        return GRIDS[world]
    p = int(np.sqrt(world))
    while world % p:
        p -= 1
    return p, world // p


def plan(P, Q, rank, N, nb, k):
    """Host-only: what `rank` does at step k of the sweep (gpc_dist_plan) -> (dict, producers)."""
    out = (C.c_int * 12)()
    NBt = (N + nb - 1) // nb
    prod = (C.c_int * NBt)()
    check(lib().gpc_dist_plan(P, Q, rank, N, nb, k, out, prod))
    names = ["nbt", "local_rows", "local_cols", "col_owner", "row_owner", "diag_owner_rank", "col_first_local",
             "col_first_slot", "row_count", "la_col_first_local", "la_row_count", "bulk_skip"]
    return dict(zip(names, list(out))), list(prod)


class DistGp:
    """Sharded logLik + gradient evaluation of an FTC GP (the quantities of gpc_eval / CGp.logLikelihoodGradient)."""

    def __init__(self, kern, X, y, bias=None, scale=None, grid=None, nb=1024, backend="nccl", devices=None, device=None,
                 group=None):
        X = fmat(X)
        y = fmat(np.asarray(y, dtype=np.float64).reshape(X.shape[0], -1))
        self.kern = kern
        self.N, self.D = X.shape
        self.d = y.shape[1]
        bias = np.zeros(self.d) if bias is None else np.asarray(bias, dtype=np.float64).reshape(self.d)
        scale = np.ones(self.d) if scale is None else np.asarray(scale, dtype=np.float64).reshape(self.d)
        self.m = fmat((y - bias) / scale)   # CGp::updateM, CGp.cpp:248-260
        self.nb = int(nb)
        self._h = C.c_void_p()
        L = lib()
        if backend == "local":
            devices = list(devices if devices is not None else [0])
            self.world, self.rank = len(devices), 0
            P, Q = grid or default_grid(self.world)
            arr = (C.c_int * len(devices))(*devices)
            check(L.gpc_dist_create_local(C.byref(self._h), arr, len(devices), P, Q, self.N, self.D, self.d, self.nb))
        elif backend == "nccl":
            import torch
            import torch.distributed as dist
            if dist.is_initialized():
                self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
            else:
                self.world, self.rank = 1, 0
            P, Q = grid or default_grid(self.world)
            idbuf = (C.c_ubyte * 128)()
            if self.world > 1:
                if self.rank == 0:
                    check(L.gpc_dist_unique_id(idbuf))
                obj = [bytes(idbuf)]
                dist.broadcast_object_list(obj, src=0, group=group)   # the launcher's one job: carry the id
                idbuf = (C.c_ubyte * 128).from_buffer_copy(obj[0])
            dev = torch.cuda.current_device() if device is None else device
            check(L.gpc_dist_create_nccl(C.byref(self._h), dev, self.rank, self.world, idbuf, P, Q, self.N, self.D, self.d,
                                         self.nb))
        else:
            raise ValueError("backend must be 'nccl' or 'local'")
        self.P, self.Q = P, Q
        self.backend = backend
        self._X = X
        self.set_data()
        self.jitter = 0.0

    def set_data(self):
        """(re-)upload X and m from the host to every rank"""
        check(lib().gpc_dist_set_data(self._h, ptr(self._X), self.N, ptr(self.m), self.N))

    def close(self):
        if self._h:
            lib().gpc_dist_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def evaluate(self):
        """(logdet, quad, g_natural); raises MatrixNonPosDef like CMatrix::jitChol when the jitter schedule gives up."""
        arr, n, keep = self.kern._kcomps()
        out = np.zeros(3)
        g = np.zeros(self.kern.getNumParams())
        rc = check(lib().gpc_dist_eval(self._h, arr, n, ptr(out), ptr(g)))
        if rc > 0:
            raise _lib.MatrixNonPosDef(rc)
        self.jitter = float(out[2])
        return float(out[0]), float(out[1]), g

    def logLikelihoodGradient(self):
        """(g_transformed, ll) like CGp.logLikelihoodGradient (CGp.cpp:1016-1144)."""
        logdet, quad, g = self.evaluate()
        ll = -0.5 * (quad + self.d * logdet) - self.d * self.N * 0.5 * np.log(2.0 * np.pi)
        return g * self.kern._gradfacts(), ll

    def download_kinv(self):
        """the blocks of K^-1 this process holds, in an N x N matrix of zeros (local back-end: all of K^-1)"""
        out = np.zeros((self.N, self.N), order="F")
        check(lib().gpc_dist_download_kinv(self._h, ptr(out), self.N))
        return out

    def info(self):
        o = (C.c_int64 * 8)()
        ms = (C.c_double * 5)()
        check(lib().gpc_dist_info(self._h, o, ms))
        keys = ["ranks", "steps", "local_matrix_bytes", "panel_buffer_bytes", "bytes_broadcast_per_step", "launches", "nb",
                "backend"]
        d = dict(zip(keys, [int(v) for v in o]))
        d["backend"] = "nccl" if d["backend"] == 1 else "local"
        d["grid"] = [self.P, self.Q]
        d["phases_ms"] = dict(zip(["kbuild", "sweep", "alpha", "grad", "reduce"], [float(v) for v in ms]))
        return d
