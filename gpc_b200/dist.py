"""Multi-GPU exact-GP evaluation: one process per GPU, `torch.distributed` for the plumbing (NCCL on GPUs), every flop
in libgpc_b200.so through the device-level C ABI (gpc_dev_*).  SURVEY.md 8(e).

Layout: the N x N problem is split into block columns of width NB, dealt round-robin to the ranks (1-D block-cyclic).
  K build      each rank builds its own block columns (no communication; X replicated)
  potrf        right-looking with a look-ahead of one panel: the owner factors the diagonal block + solves the panel
               below (gpc_dev_potrf / gpc_dev_trsm) and BROADCASTS the panel (+ the inverses of its 128-blocks)
               asynchronously; every rank applies the rank-NB update to the block columns it owns (gpc_dev_gemm), the
               owner of the next panel to that column first.  After the loop every rank holds all of L.
  inverse      W = L^-1 by block columns: every rank solves W_J' L_sub' = [I 0] for ALL its own block columns first
               (N^3/3 flop in total, split over ranks, no communication), then the W_J' blocks are ALL-GATHERED, then
               each rank forms its own block columns of K^-1 = W'W as GEMMs (another N^3/3 split over ranks)
  alpha        two triangular solves with the replicated L (O(N^2 d), every rank)
  gradient     fused pass over the owned block columns of K^-1 (gpc_dev_grad_cols), ALL-REDUCE of P doubles;
               logdet partials all-reduced likewise
Collectives: broadcast (panels), all_gather-by-broadcast (W blocks), all_reduce (scalars).  With world_size == 1 the
same code runs without a process group.

`ops` is the compute backend: DeviceOps (CUDA, the product) below.  The CPU tests inject a numpy stand-in to check the
distributed schedule under gloo (tests/test_dist_cpu.py); there is no CPU backend in this package.
"""
import ctypes as C
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, lib


def _ptr(t, offset_elems=0):
    return C.c_void_p(t.data_ptr() + 8 * int(offset_elems))


class DeviceOps:
    """CUDA backend.  A column-major (rows x cols) matrix is a contiguous torch tensor of shape (cols, rows)."""

    def __init__(self, device):
        self.device = torch.device("cuda", device)
        self.idx = device
        torch.cuda.set_device(device)
        self._h = C.c_void_p()
        check(lib().gpc_dev_create(C.byref(self._h), device, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        self.info = torch.zeros(2, dtype=torch.int32, device=self.device)
        self.logdet = torch.zeros(1, dtype=torch.float64, device=self.device)

    def close(self):
        if self._h:
            lib().gpc_dev_destroy(self._h)
            self._h = None

    def _sync_stream(self):
        check(lib().gpc_dev_set_stream(self._h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def zeros(self, *shape):
        return torch.zeros(*shape, dtype=torch.float64, device=self.device)

    def empty(self, *shape):
        return torch.empty(*shape, dtype=torch.float64, device=self.device)

    def from_numpy(self, a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(self.device)

    def launch_count(self):
        return int(lib().gpc_dev_launch_count(self._h))

    def reset_scalars(self):
        self.info.zero_()
        self.logdet.zero_()

    # --- kernels (all on torch's current stream) -------------------------------------------------------------
    def kbuild_cols(self, kcomps, Xt, n, Lt, col0, ncols):
        arr, nc, keep = kcomps
        Np, D = Lt.shape[1], Xt.shape[0]
        self._sync_stream()
        check(lib().gpc_dev_kbuild_cols(self._h, arr, nc, _ptr(Xt), Np, n, Np, D, col0, ncols, _ptr(Lt), Np))

    def potrf_block(self, Lt, k0, nb, n, Dinv):
        Np = Lt.shape[1]
        self._sync_stream()
        check(lib().gpc_dev_potrf(self._h, _ptr(Lt, k0 + k0 * Np), Np, nb, k0, n, _ptr(Dinv, k0 * 128),
                                  _ptr(self.info), _ptr(self.logdet)))

    def trsm_panel(self, Lt, k0, nb, Dinv):
        Np = Lt.shape[1]
        m = Np - k0 - nb
        if m <= 0:
            return
        self._sync_stream()
        check(lib().gpc_dev_trsm(self._h, b"T", _ptr(Lt, (k0 + nb) + k0 * Np), Np, m, _ptr(Lt, k0 + k0 * Np), Np, nb,
                                 _ptr(Dinv, k0 * 128)))

    def update_cols(self, Lt, j0, nbj, k0, nbk):
        """L[j0:, j0:j0+nbj] -= L[j0:, k0:k0+nbk] L[j0:j0+nbj, k0:k0+nbk]'"""
        Np = Lt.shape[1]
        self._sync_stream()
        check(lib().gpc_dev_gemm(self._h, 0, 0, 0, Np - j0, nbj, nbk, -1.0, _ptr(Lt, j0 + k0 * Np), Np,
                                 _ptr(Lt, j0 + k0 * Np), Np, 1.0, _ptr(Lt, j0 + j0 * Np), Np))

    def winv_block(self, Lt, j0, nb, Dinv):
        """Block column J of W = L^-1, returned as a contiguous tensor (nb, Np-j0): row a holds W[j0:, j0+a].
        Computed transposed as a right-sided solve X L_sub' = [I 0] with L_sub = L[j0:, j0:]."""
        Np = Lt.shape[1]
        WTj = self.zeros(Np - j0, nb)    # column-major nb x (Np-j0), ld nb
        WTj[:nb, :].fill_diagonal_(1.0)
        self._sync_stream()
        check(lib().gpc_dev_trsm(self._h, b"T", _ptr(WTj), nb, nb, _ptr(Lt, j0 + j0 * Np), Np, Np - j0,
                                 _ptr(Dinv, j0 * 128)))
        return WTj.t().contiguous()

    def kinv_cols(self, Kc, Wc, j0, nb, jl):
        """Kc[j0:, jl:jl+nb] = (W[:, j0:])' W[:, j0:j0+nb]  (rows i >= j0 of block column J of K^-1 = W'W): one GEMM
        whose A operand is upper triangular (W[k, i] = 0 for k < i), so every row tile skips its zero k range."""
        Np = Kc.shape[1]
        m = Np - j0
        self._sync_stream()
        check(lib().gpc_dev_gemm(self._h, 1, 1, 2, m, nb, m, 1.0, _ptr(Wc, j0 + j0 * Np), Np, _ptr(Wc, j0 + j0 * Np), Np,
                                 0.0, _ptr(Kc, j0 + jl * Np), Np))

    def alpha_solve(self, Lt, Dinv, mt):
        """alpha = L^-T L^-1 m with the replicated factor; mt is (d, Np).  Returns alpha_t (d, Np)."""
        Np, d = Lt.shape[1], mt.shape[0]
        dp = 128 * ((d + 127) // 128)
        T = self.zeros(Np, dp)      # the d x Np matrix m' padded to dp rows, column-major with ld dp
        T[:, :d] = mt.t()
        self._sync_stream()
        check(lib().gpc_dev_trsm(self._h, b"T", _ptr(T), dp, dp, _ptr(Lt), Np, Np, _ptr(Dinv)))
        check(lib().gpc_dev_trsm(self._h, b"N", _ptr(T), dp, dp, _ptr(Lt), Np, Np, _ptr(Dinv)))
        return T[:, :d].t().contiguous()

    def grad_cols(self, kcomps, Xt, n, Kc, col0, ncols, jl, alpha_t):
        arr, nc, keep = kcomps
        Np, D, d = Kc.shape[1], Xt.shape[0], alpha_t.shape[0]
        P = sum(arr[i].nparams for i in range(nc))
        g = np.zeros(P)
        self._sync_stream()
        # Cg must address the full matrix: column col0 of K^-1 lives at local column jl of Kc
        base = Kc.data_ptr() + 8 * (jl - col0) * Np
        check(lib().gpc_dev_grad_cols(self._h, arr, nc, _ptr(Xt), Np, n, D, col0, ncols, C.c_void_p(base), Np,
                                      _ptr(alpha_t), Np, d, g.ctypes.data_as(C.c_void_p)))
        return g


class DistGp:
    """Sharded logLik + gradient evaluation of an FTC GP (same quantities as gpc_eval / CGp.logLikelihoodGradient)."""

    def __init__(self, ops, kern, X, m, NB=1024, group=None):
        self.ops, self.kern, self.group = ops, kern, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        X = np.asarray(X, dtype=np.float64)
        m = np.asarray(m, dtype=np.float64).reshape(X.shape[0], -1)
        self.N, self.D = X.shape
        self.d = m.shape[1]
        assert NB % 128 == 0
        self.NB = NB
        self.Np = NB * ((self.N + NB - 1) // NB)
        self.nblk = self.Np // NB
        Np = self.Np
        Xp = np.zeros((self.D, Np))
        Xp[:, :self.N] = X.T
        mp = np.zeros((self.d, Np))
        mp[:, :self.N] = m.T
        self.Xt = ops.from_numpy(Xp)     # column-major Np x D
        self.mt = ops.from_numpy(mp)     # column-major Np x d
        self.Lt = ops.empty(Np, Np)      # column-major Np x Np: own K columns, then the replicated factor
        self.Dinv = ops.zeros(Np // 128, 128, 128)
        self.owned = [b for b in range(self.nblk) if b % self.world == self.rank]
        self.Kc = ops.empty(len(self.owned) * NB, Np)  # own block columns of K^-1 (column-major Np x ncols)
        self.Wc = ops.zeros(Np, Np)      # W = L^-1, column-major (replicated after the all-gather; zero above the diagonal)
        self.times = {}

    def owner(self, b):
        return b % self.world

    def _bcast(self, t, src):
        if self.world > 1:
            dist.broadcast(t, src=src, group=self.group)

    def _allreduce(self, t, op=None):
        if self.world > 1:
            dist.all_reduce(t, op=op or dist.ReduceOp.SUM, group=self.group)

    def _tick(self, name):
        """phase timing (GPC_DIST_TIMING=1): synchronises, so only for diagnosis"""
        if not self._timing:
            return
        if self.Lt.is_cuda:
            torch.cuda.synchronize()
        now = time.time()
        self.times[name] = self.times.get(name, 0.0) + (now - self._t0)
        self._t0 = now

    def evaluate(self):
        """returns (logdet, quad, g_natural)."""
        ops, NB, Np, nblk = self.ops, self.NB, self.Np, self.nblk
        kc = self.kern._kcomps()
        ops.reset_scalars()
        self._timing = bool(os.environ.get("GPC_DIST_TIMING"))
        self.times = {}
        if self._timing and self.Lt.is_cuda:
            torch.cuda.synchronize()
        self._t0 = time.time()
        # ---- K build: own block columns, straight into the factor buffer
        for b in self.owned:
            ops.kbuild_cols(kc, self.Xt, self.N, self.Lt, b * NB, NB)
        self._tick("kbuild")
        # ---- right-looking Cholesky with panel broadcasts and a look-ahead of one panel: the owner of panel kb+1
        #      updates that block column first, factors it and starts its (asynchronous) broadcast while everybody is
        #      still applying panel kb to the rest of their columns
        def start_panel(kb):
            k0 = kb * NB
            src = self.owner(kb)
            if self.rank == src:
                ops.potrf_block(self.Lt, k0, NB, self.N, self.Dinv.view(-1))
                ops.trsm_panel(self.Lt, k0, NB, self.Dinv.view(-1))
            if self.world == 1:
                return None, None, []
            panel = self.Lt[k0:k0 + NB, k0:]            # nb block-column, rows k0.. (strided view)
            buf = panel.contiguous() if self.rank == src else ops.empty(NB, Np - k0)
            dblk = self.Dinv[k0 // 128:(k0 + NB) // 128]
            works = [dist.broadcast(buf, src=src, group=self.group, async_op=True),
                     dist.broadcast(dblk, src=src, group=self.group, async_op=True)]
            return panel, buf, works

        pending = start_panel(0)
        self._tick("potrf_panel")
        for kb in range(nblk):
            k0 = kb * NB
            panel, buf, works = pending
            for w in works:
                w.wait()
            if works and self.rank != self.owner(kb):
                panel.copy_(buf)
            del buf, panel
            self._tick("potrf_bcast")
            if kb + 1 < nblk:
                if self.rank == self.owner(kb + 1):
                    ops.update_cols(self.Lt, (kb + 1) * NB, NB, k0, NB)
                pending = start_panel(kb + 1)
                self._tick("potrf_panel")
            for b in self.owned:
                if b > kb + 1:
                    ops.update_cols(self.Lt, b * NB, NB, k0, NB)
            self._tick("potrf_update")
        # ---- status: first non-positive pivot (max over ranks of a "first or zero" is good enough to fail loudly)
        st = torch.zeros(2, dtype=torch.float64, device=self.Lt.device)
        st[0] = ops.info[0].double()
        st[1] = ops.logdet[0]
        info_t = st[:1].clone()
        self._allreduce(info_t, dist.ReduceOp.MAX if self.world > 1 else None)
        if float(info_t.item()) != 0.0:
            raise _lib.MatrixNonPosDef(int(info_t.item()))
        ld_t = st[1:].clone()
        self._allreduce(ld_t)
        logdet = float(ld_t.item())
        # ---- W = L^-1 by block columns.  Every rank first solves ALL the block columns it owns (no communication:
        #      the ranks work concurrently), only then are the compact non-zero parts all-gathered by broadcasts.
        #      (Solving and broadcasting block by block serialises the ranks: the owner of block b+1 sits in the
        #      broadcast of block b while its peer is still solving.)
        mine = {}
        for b in self.owned:
            mine[b] = ops.winv_block(self.Lt, b * NB, NB, self.Dinv.view(-1))
        self._tick("winv_solve")
        for b in range(nblk):
            j0 = b * NB
            src = self.owner(b)
            buf = mine.pop(b) if self.rank == src else ops.empty(NB, Np - j0)
            self._bcast(buf, src)
            self.Wc[j0:j0 + NB, j0:].copy_(buf)
            del buf
        self._tick("winv_bcast")
        # ---- own block columns of K^-1 = W'W (rows i >= j0)
        for jl, b in enumerate(self.owned):
            ops.kinv_cols(self.Kc, self.Wc, b * NB, NB, jl * NB)
        self._tick("kinv_gemm")
        # ---- alpha (replicated), quadratic form
        alpha_t = ops.alpha_solve(self.Lt, self.Dinv.view(-1), self.mt)
        quad = float((alpha_t * self.mt).sum().item())
        self._tick("alpha")
        # ---- gradient partial sums over the owned columns
        g = None
        for jl, b in enumerate(self.owned):
            gb = ops.grad_cols(kc, self.Xt, self.N, self.Kc, b * NB, NB, jl * NB, alpha_t)
            g = gb if g is None else g + gb
        if g is None:
            g = np.zeros(self.kern.getNumParams())
        gt = torch.from_numpy(g).to(self.Lt.device)
        self._allreduce(gt)
        self._tick("grad")
        return logdet, quad, gt.cpu().numpy()

    def logLikelihoodGradient(self):
        """(g_transformed, ll) like CGp.logLikelihoodGradient."""
        logdet, quad, g = self.evaluate()
        ll = -0.5 * (quad + self.d * logdet) - self.d * self.N * 0.5 * np.log(2.0 * np.pi)
        return g * self.kern._gradfacts(), ll
