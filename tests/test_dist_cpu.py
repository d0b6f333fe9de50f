"""Multi-rank path on CPU (gloo, world sizes 2, 4 and 8): checks the distributed SCHEDULE of gpc_b200/dist.py -- block-cyclic
ownership, panel broadcasts, the all-gather of the W blocks, the all-reduces -- with a numpy stand-in for the device
kernels (test double defined here; the package has no CPU backend).  The result must equal the oracle's ll/gradient."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class FakeOps:
    """numpy/torch-CPU stand-in for gpc_b200.dist.DeviceOps with the same layout conventions (test double)."""

    def __init__(self, types, D):
        from oracle import gp_oracle as O
        self.O, self.types, self.D = O, types, D
        self.info = torch.zeros(2, dtype=torch.int32)
        self.logdet = torch.zeros(1, dtype=torch.float64)

    def zeros(self, *s):
        return torch.zeros(*s, dtype=torch.float64)

    def empty(self, *s):
        return torch.full(s, float("nan"), dtype=torch.float64)

    def from_numpy(self, a):
        return torch.from_numpy(np.ascontiguousarray(a))

    def reset_scalars(self):
        self.info.zero_()
        self.logdet.zero_()

    def _kern(self, kcomps):
        arr, nc, keep = kcomps
        return [(self.types[i], np.array(keep[i])) for i in range(nc)]

    def kbuild_cols(self, kcomps, Xt, n, Lt, col0, ncols):
        Np = Lt.shape[1]
        X = Xt.numpy().T[:n]
        K = np.eye(Np)
        K[:n, :n] = self.O.kern_compute(self._kern(kcomps), X)
        Lt[col0:col0 + ncols, :] = torch.from_numpy(K[:, col0:col0 + ncols].T.copy())

    def potrf_block(self, Lt, k0, nb, n, Dinv):
        A = Lt[k0:k0 + nb, k0:k0 + nb].numpy().T
        A = np.tril(A) + np.tril(A, -1).T
        L, info = self.O.chol_lower(A)
        if info and self.info[0] == 0:
            self.info[0] = k0 + info
        Lt[k0:k0 + nb, k0:k0 + nb] = torch.from_numpy(np.tril(L).T.copy())   # upper part: zeros (device leaves junk)
        nv = max(0, min(nb, n - k0))
        self.logdet += 2.0 * float(np.sum(np.log(np.diag(L)[:nv])))
        D3 = Dinv.view(-1, 128, 128)
        for b in range(nb // 128):
            blk = L[b * 128:(b + 1) * 128, b * 128:(b + 1) * 128]
            D3[k0 // 128 + b] = torch.from_numpy(np.linalg.inv(blk).T.copy())  # column-major 128x128

    def trsm_panel(self, Lt, k0, nb, Dinv):
        if Lt.shape[1] - k0 - nb <= 0:
            return
        L = np.tril(Lt[k0:k0 + nb, k0:k0 + nb].numpy().T)
        B = Lt[k0:k0 + nb, k0 + nb:].numpy().T            # rows below x nb
        X = np.linalg.solve(L, B.T).T                     # X L' = B
        Lt[k0:k0 + nb, k0 + nb:] = torch.from_numpy(X.T.copy())

    def update_cols(self, Lt, j0, nbj, k0, nbk):
        A = Lt[k0:k0 + nbk, j0:].numpy().T                # L[j0:, k0:k0+nbk]
        Bm = A[:nbj]
        C_ = Lt[j0:j0 + nbj, j0:].numpy().T
        Lt[j0:j0 + nbj, j0:] = torch.from_numpy((C_ - A @ Bm.T).T.copy())

    def winv_block(self, Lt, j0, nb, Dinv):
        Lsub = np.tril(Lt[j0:, j0:].numpy().T)
        E = np.zeros((Lsub.shape[0], nb))
        E[:nb] = np.eye(nb)
        import scipy.linalg as sla
        W = sla.solve_triangular(Lsub, E, lower=True)     # (Np-j0) x nb = W[j0:, J]; exact zeros above the diagonal
        return torch.from_numpy(W.T.copy())               # (nb, Np-j0): row a = W[j0:, j0+a]

    def kinv_cols(self, Kc, Wc, j0, nb, jl):
        W = Wc.numpy().T                                  # W[k, i]
        assert np.all(np.triu(W, 1) == 0.0)               # the schedule must leave W lower triangular
        blk = W[j0:, j0:].T @ W[j0:, j0:j0 + nb]          # (Np-j0) x nb
        Kc[jl:jl + nb, j0:] = torch.from_numpy(blk.T.copy())

    def alpha_solve(self, Lt, Dinv, mt):
        L = np.tril(Lt.numpy().T)
        m = mt.numpy().T
        a = np.linalg.solve(L.T, np.linalg.solve(L, m))
        return torch.from_numpy(a.T.copy())

    def grad_cols(self, kcomps, Xt, n, Kc, col0, ncols, jl, alpha_t):
        Np = Kc.shape[1]
        X = Xt.numpy().T[:n]
        alpha = alpha_t.numpy().T[:n]
        d = alpha.shape[1]
        cols = Kc[jl:jl + ncols].numpy().T                # Np x ncols: K^-1[:, col0:col0+ncols] (rows >= col valid)
        cg = np.zeros((n, n))
        for jj in range(ncols):
            j = col0 + jj
            if j >= n:
                break
            v = -0.5 * (d * cols[j:n, jj] - alpha[j:n] @ alpha[j])
            cg[j:n, j] = v
            cg[j, j:n] = v
        return self.O.kern_grad_params(self._kern(kcomps), X, cg)


def _worker(rank, world, port, N, D, NB, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    import gpc_b200 as G
    from gpc_b200.dist import DistGp
    rng = np.random.default_rng(7)
    X = rng.standard_normal((N, D))
    y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((N, 2))
    types = ["rbf", "matern32", "white"]
    kern = G.make_kern(types, D, [-0.7, 0.2, 0.4, -0.3, -2.0])
    gp = DistGp(FakeOps(types, D), kern, X, y, NB=NB)
    g, ll = gp.logLikelihoodGradient()
    if rank == 0:
        np.savez(out, g=g, ll=ll, owned=np.array(gp.owned))
    if world > 1:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,N,NB", [(1, 200, 128), (2, 300, 128), (2, 700, 256), (4, 900, 128), (8, 1100, 128),
                                         (8, 300, 128)])   # the last: more ranks than block columns
def test_distributed_schedule_matches_oracle(tmp_path, world, N, NB):
    from oracle import gp_oracle as O
    D = 3
    out = str(tmp_path / "res.npz")
    port = 29500 + (os.getpid() % 2000)
    if world == 1:
        _worker(0, 1, port, N, D, NB, out)
    else:
        mp.spawn(_worker, args=(world, port, N, D, NB, out), nprocs=world, join=True)
    r = np.load(out)
    rng = np.random.default_rng(7)
    X = rng.standard_normal((N, D))
    y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((N, 2))
    ref = O.gp_loglik_grad(O.kern_from_trans(["rbf", "matern32", "white"], [-0.7, 0.2, 0.4, -0.3, -2.0], D), X, y)
    assert abs(float(r["ll"]) - ref["ll"]) <= 1e-8 * max(1.0, abs(ref["ll"]))
    assert np.max(np.abs(r["g"] - ref["g"]) / np.maximum(1.0, np.abs(ref["g"]))) < 1e-8


def test_block_cyclic_ownership():
    import gpc_b200 as G
    from gpc_b200.dist import DistGp
    kern = G.make_kern(["rbf", "white"], 2)
    gp = DistGp(FakeOps(["rbf", "white"], 2), kern, np.zeros((1000, 2)), np.zeros((1000, 1)), NB=256)
    assert gp.Np == 1024 and gp.nblk == 4 and gp.owned == [0, 1, 2, 3]
    assert [gp.owner(b) for b in range(4)] == [0, 0, 0, 0]
