"""Multi-rank path on CPU.  The distributed evaluation itself (gpc_b200/csrc/dist.cu) needs GPUs; what runs here is
(1) the algorithm -- the fused one-sweep K -> K^-1 on a 2-D block-cyclic layout -- restated in numpy with one local
    matrix per rank and explicit slot broadcasts, driven by the ownership / schedule helpers OF THE LIBRARY
    (gpc_dist_plan, host-only): in one process for several grids, and as real processes under gloo (world size 2 and 4);
(2) the schedule invariants the device code relies on (every slot has exactly one producer, look-ahead strips + bulk cover
    every block exactly once per step, ...).
The numpy restatement is a test double defined here; the package has no CPU backend."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpc_b200.dist import default_grid, plan  # noqa: E402


class RankModel:
    """one rank of the P x Q grid: local blocks as a dict {(i, j): nb x nb array}, blocks on / below the diagonal only"""

    def __init__(self, K, nb, P, Q, rank):
        self.nb, self.P, self.Q, self.rank = nb, P, Q, rank
        self.p, self.q = rank // Q, rank % Q
        self.N = K.shape[0]
        self.NBt = self.N // nb
        self.T = {}
        for i in range(self.NBt):
            for j in range(i + 1):
                if i % P == self.p and j % Q == self.q:
                    self.T[(i, j)] = K[i * nb:(i + 1) * nb, j * nb:(j + 1) * nb].copy()
        self.logdet = 0.0

    def factor(self, k):
        A = self.T[(k, k)]
        L = np.linalg.cholesky(np.tril(A) + np.tril(A, -1).T)
        self.logdet += 2.0 * np.log(np.diag(L)).sum()
        return np.linalg.inv(L)

    def produce(self, k, W):
        """this rank's slots of panel k, following gpc_dist_plan"""
        pl, _ = plan(self.P, self.Q, self.rank, self.N, self.nb, k)
        slots = {}
        if pl["col_owner"]:
            il = pl["col_first_local"]
            g = pl["col_first_slot"]
            assert g == il * self.P + self.p
            while g < self.NBt:
                assert g > k
                slots[g] = self.T[(g, k)] @ W.T
                g += self.P
        if pl["row_owner"]:
            for jl in range(pl["row_count"]):
                g = jl * self.Q + self.q
                assert g < k
                slots[g] = (W @ self.T[(k, g)]).T
        if pl["diag_owner_rank"] == self.rank:
            slots[k] = W.T.copy()
        return slots

    def update(self, k, S, only=None, skip=None):
        """T_ij <- beta T_ij - sgn_i S_i S_j'; only / skip: a block row+column index"""
        done = []
        for (i, j) in self.T:
            in_strip = (i == only or j == only) if only is not None else True
            if not in_strip or (skip is not None and (i == skip or j == skip)):
                continue
            sgn = 1.0 if i > k else -1.0
            beta = 0.0 if (i == k or j == k) else 1.0
            self.T[(i, j)] = beta * self.T[(i, j)] - sgn * S[i] @ S[j].T
            done.append((i, j))
        return done


def assemble(models, N, nb):
    out = np.zeros((N, N))
    for m in models:
        for (i, j), blk in m.T.items():
            if i == j:
                blk = np.tril(blk) + np.tril(blk, -1).T
            out[i * nb:(i + 1) * nb, j * nb:(j + 1) * nb] = blk
            out[j * nb:(j + 1) * nb, i * nb:(i + 1) * nb] = blk.T
    return out


def spd(n, seed=0):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n))
    return A @ A.T + n * np.eye(n)


@pytest.mark.parametrize("P,Q,NBt", [(1, 1, 5), (1, 2, 6), (2, 2, 7), (2, 4, 9), (4, 2, 6), (3, 1, 5), (2, 4, 3)])
def test_one_sweep_schedule_single_process(P, Q, NBt):
    """all ranks in one process, with the look-ahead order of the device code: strips of step k-1 first, then panel k,
    then the bulk of step k without block row / column k+1.  Every block must be updated exactly once per step."""
    nb = 8
    N = NBt * nb
    K = spd(N, seed=P * 10 + Q)
    world = P * Q
    models = [RankModel(K, nb, P, Q, r) for r in range(world)]
    Sprev = None
    for k in range(NBt):
        touched = []
        if k >= 1:  # look-ahead strips: panel k-1 on block row / column k
            for m in models:
                touched += [("la", b) for b in m.update(k - 1, Sprev, only=k)]
        pl0, producers = plan(P, Q, 0, N, nb, k)
        assert pl0["nbt"] == NBt and len(producers) == NBt
        W = models[pl0["diag_owner_rank"]].factor(k)
        S = {}
        for m in models:
            mine = m.produce(k, W)
            for g, blk in mine.items():
                assert producers[g] == m.rank and g not in S   # exactly one producer per slot, as the plan says
                S[g] = blk
        assert sorted(S) == list(range(NBt))
        skip = pl0["bulk_skip"]
        assert skip == (k + 1 if k + 1 < NBt else -1)
        for m in models:
            m.update(k, S, skip=skip if skip >= 0 else None)
        Sprev = S
    Kinv = assemble(models, N, nb)
    ref = np.linalg.inv(K)
    assert np.max(np.abs(Kinv - ref)) < 1e-12 * np.max(np.abs(ref)) * N
    logdet = sum(m.logdet for m in models)
    assert logdet == pytest.approx(np.linalg.slogdet(K)[1], rel=1e-13)


def test_plan_local_extents_and_padding():
    """local block counts add up, ragged N is padded to whole blocks, more ranks than block rows is fine"""
    for (P, Q, N, nb) in [(2, 4, 1000, 128), (2, 2, 129, 128), (2, 4, 300, 128), (1, 2, 65536, 1024), (4, 2, 5000, 256)]:
        NBt = (N + nb - 1) // nb
        rows = [plan(P, Q, p * Q, N, nb, 0)[0]["local_rows"] for p in range(P)]
        cols = [plan(P, Q, q, N, nb, 0)[0]["local_cols"] for q in range(Q)]
        assert sum(rows) == NBt and sum(cols) == NBt
        for k in range(0, NBt, max(1, NBt // 5)):
            seen = {}
            for r in range(P * Q):
                pl, prod = plan(P, Q, r, N, nb, k)
                assert pl["diag_owner_rank"] == (k % P) * Q + (k % Q)
                assert prod[k] == pl["diag_owner_rank"]
                seen[r] = pl
            assert sum(1 for r in seen if seen[r]["col_owner"]) == P
            assert sum(1 for r in seen if seen[r]["row_owner"]) == Q
    assert default_grid(8) == (2, 4) and default_grid(4) == (2, 2) and default_grid(2) == (1, 2) and default_grid(1) == (1, 1)


def _gloo_worker(rank, world, port, P, Q, NBt, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nb = 8
    N = NBt * nb
    K = spd(N, seed=7)
    m = RankModel(K, nb, P, Q, rank)
    Sprev = None
    for k in range(NBt):
        if k >= 1:
            m.update(k - 1, Sprev, only=k)
        pl, producers = plan(P, Q, rank, N, nb, k)
        # W_kk from the owner of block (k, k)
        W = torch.zeros(nb, nb, dtype=torch.float64)
        if pl["diag_owner_rank"] == rank:
            W = torch.from_numpy(m.factor(k))
        dist.broadcast(W, src=pl["diag_owner_rank"])
        mine = m.produce(k, W.numpy())
        # one broadcast per slot, from the rank that produced it (what the device code does with ncclBroadcast)
        S = {}
        for g in range(NBt):
            t = torch.from_numpy(np.ascontiguousarray(mine[g])) if producers[g] == rank else torch.zeros(nb, nb, dtype=torch.float64)
            assert (g in mine) == (producers[g] == rank)
            dist.broadcast(t, src=producers[g])
            S[g] = t.numpy().copy()
        skip = pl["bulk_skip"]
        m.update(k, S, skip=skip if skip >= 0 else None)
        Sprev = S
    ld = torch.tensor([m.logdet], dtype=torch.float64)
    dist.all_reduce(ld)
    objs = [None] * world
    dist.all_gather_object(objs, {kk: v for kk, v in m.T.items()})
    if rank == 0:
        Kinv = np.zeros((N, N))
        for T in objs:
            for (i, j), blk in T.items():
                if i == j:
                    blk = np.tril(blk) + np.tril(blk, -1).T
                Kinv[i * nb:(i + 1) * nb, j * nb:(j + 1) * nb] = blk
                Kinv[j * nb:(j + 1) * nb, i * nb:(i + 1) * nb] = blk.T
        np.savez(out, Kinv=Kinv, logdet=float(ld.item()), K=K)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,NBt", [(2, 5), (4, 6)])
def test_one_sweep_schedule_gloo(tmp_path, world, NBt):
    P, Q = default_grid(world)
    out = str(tmp_path / "res.npz")
    mp.spawn(_gloo_worker, args=(world, 29700 + os.getpid() % 500 + world, P, Q, NBt, out), nprocs=world, join=True)
    z = np.load(out)
    ref = np.linalg.inv(z["K"])
    assert np.max(np.abs(z["Kinv"] - ref)) < 1e-12 * np.max(np.abs(ref)) * z["K"].shape[0]
    assert float(z["logdet"]) == pytest.approx(np.linalg.slogdet(z["K"])[1], rel=1e-13)
