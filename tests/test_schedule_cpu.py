"""CPU tests of host-side scheduling logic that needs no device: the launch boundaries of a wave-limited product on the
tensor-core engine (GemmCall::sm_limit, gpc_b200/csrc/ozaki.cu: oz_wave_cuts), checked against a restatement of the
kernel's own tile raster (oz_gemm_kernel: groups of 8 row tiles, column by column; lower mode keeps bn <= 2 bm + 1)."""
import ctypes as C

import pytest

from gpc_b200._lib import lib


def _raster(t, tiles_m, tiles_n):
    GM = 8
    per_group = GM * tiles_n
    grp, rem = divmod(t, per_group)
    first = grp * GM
    gsz = min(GM, tiles_m - first)
    return first + rem % gsz, rem // gsz


def _cuts(tm, tn, lower, limit):
    L = lib()
    L.gpc_oz_wave_cuts.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong), C.c_int]
    n = L.gpc_oz_wave_cuts(tm, tn, lower, limit, None, 0)
    assert n >= 2
    buf = (C.c_longlong * n)()
    assert L.gpc_oz_wave_cuts(tm, tn, lower, limit, buf, n) == n
    return list(buf)


@pytest.mark.parametrize("tm,tn,lower", [(32, 64, 0), (32, 64, 1), (64, 128, 1), (7, 14, 1), (9, 5, 0), (1, 2, 1), (33, 32, 0)])
@pytest.mark.parametrize("limit", [0, 16, 100, 132, 148, 100000])
def test_wave_cuts_cover_every_tile_once_and_respect_the_limit(tm, tn, lower, limit):
    cuts = _cuts(tm, tn, lower, limit)
    ntiles = tm * tn
    assert cuts[0] == 0 and cuts[-1] == ntiles
    assert all(b >= a for a, b in zip(cuts, cuts[1:]))
    if limit <= 0 or ntiles <= limit:
        assert cuts == [0, ntiles]
        return
    seen = set()
    for a, b in zip(cuts, cuts[1:]):
        live = 0
        for t in range(a, b):
            bm, bn = _raster(t, tm, tn)
            assert 0 <= bm < tm and 0 <= bn < tn
            assert (bm, bn) not in seen
            seen.add((bm, bn))
            if not lower or bn <= 2 * bm + 1:
                live += 1
        assert live <= max(limit, 8), (a, b, live)   # a single column of a group (<= 8 tiles) is never split
    assert len(seen) == ntiles


def test_wave_cuts_rejects_empty_shapes():
    assert lib().gpc_oz_wave_cuts(0, 4, 0, 10, None, 0) < 0


# ---------------------------------------------------------------------------------------------------------
# numpy restatements of device schedules (documentation that runs): the same index arithmetic as the CUDA code,
# checked against numpy's own Cholesky / inverse
import numpy as np  # noqa: E402


def _spd(n, seed):
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((n, n))
    return M @ M.T + n * np.eye(n)


def test_diagonal_block_schedule_restated():
    """potrf_leaf_kernel, phase 1 (dense.cu): per 16-column panel, warp w holds the 16 rows of the diagonal block on lanes
    0..15 and the rows j0 + 16 + 16 w .. + 15 on lanes 16..31; one pivot loop factors the block and solves the rows below
    it; then the rank-16 update of the trailing lower 8 x 8 tiles in groups of up to 4 row tiles per column tile."""
    TILE, PB = 128, 16
    A = _spd(TILE, 1)
    sA = np.tril(A).copy()

    def panel(src, dst, j0, grp, leader):
        rows = [(j0 + l) if l < PB else (j0 + PB + PB * grp + (l - PB)) for l in range(32)]
        valid = [r < TILE for r in rows]
        a = np.array([src[r, j0:j0 + PB] if v else np.zeros(PB) for r, v in zip(rows, valid)])
        piv = a[0, 0]
        for c in range(PB):
            inv = 1.0 / np.sqrt(piv)
            lc = a[:, c] * inv
            lc[c] = piv * inv
            a[:, c] = lc
            if c + 1 < PB:
                piv = a[c + 1, c + 1] - lc[c + 1] * lc[c + 1]   # broadcast before the general update below
            for c2 in range(c + 1, PB):
                a[:, c2] -= lc * lc[c2]
        for l in range(32):
            if l < PB:
                if leader:
                    dst[rows[l], j0:j0 + l + 1] = a[l, :l + 1]
            elif valid[l]:
                dst[rows[l], j0:j0 + PB] = a[l]

    def rank16(j0, I0, ng, C0):
        P = sA[:, j0:j0 + PB].copy()
        for g in range(ng):
            I = I0 + 8 * g
            sA[I:I + 8, C0:C0 + 8] -= P[I:I + 8] @ P[C0:C0 + 8].T

    for j0 in range(0, TILE, PB):
        out = sA.copy()
        for w in range(TILE // PB):
            if w == 0 or j0 + PB + PB * w < TILE:
                panel(sA, out, j0, w, w == 0)
        sA = out
        R0 = j0 + PB
        T = (TILE - R0) // 8
        for tc in range(T):
            for ti in range(tc, T, 4):
                rank16(j0, R0 + 8 * ti, min(4, T - ti), R0 + 8 * tc)
    assert np.abs(np.tril(sA) - np.linalg.cholesky(A)).max() < 1e-12


def test_top_level_lookahead_and_row_block_pipeline_restated():
    """api.cu, potrf_inv_rec at the top level: with A11 split into diagonal nodes a, b and A22 into a', b',
    TopFront forms L21 column block by column block and TopPipe forms W21 and K^-1 = W'W row block by row block;
    both must reproduce the plain block formulas."""
    n1, n2, h, h1 = 12, 10, 5, 4
    A = _spd(n1 + n2, 2)
    L = np.linalg.cholesky(A)
    W = np.linalg.inv(L)
    W11, L21ref = W[:n1, :n1], L[n1:, :n1]
    A21 = A[n1:, :n1]
    # ---- TopFront: X1 / X2 after node a, X3 / X4 after node b
    L21 = np.zeros_like(A21)
    L21[:, :h] = A21[:, :h] @ W11[:h, :h].T                                             # X1
    A22 = A[n1:, n1:] - L21[:, :h] @ L21[:, :h].T                                         # X2
    L21[:, h:] = A21[:, :h] @ W11[h:, :h].T + A21[:, h:] @ W11[h:, h:].T                  # X3
    A22 -= L21[:, h:] @ L21[:, h:].T                                                       # X4
    assert np.abs(L21 - L21ref).max() < 1e-12
    assert np.abs(np.tril(A22) - np.tril(L[n1:, n1:] @ L[n1:, n1:].T)).max() < 1e-10
    # ---- TopPipe: T = L21 W11; rows a' of W21 and their share of K^-1 early, rows b' in the tail
    T = L21 @ W11
    W22 = W[n1:, n1:]
    Kinv = np.zeros_like(A)
    Kinv[:n1, :n1] = W11.T @ W11                                                           # after node(A11)
    W21 = np.zeros((n2, n1))
    Waa = W22[:h1, :h1]
    W21[:h1] = -Waa @ T[:h1]                                                               # Y1
    Kinv[:n1, :n1] += W21[:h1].T @ W21[:h1]                                                # Y3a
    Kinv[n1:n1 + h1, :n1] = Waa.T @ W21[:h1]                                               # Y3b
    Kinv[n1:n1 + h1, n1:n1 + h1] = Waa.T @ Waa                                             # Y3c
    W21[h1:] = -(W22[h1:, :h1] @ T[:h1] + W22[h1:, h1:] @ T[h1:])                          # Z1
    assert np.abs(W21 - W[n1:, :n1]).max() < 1e-12
    m = n1 + h1
    R = np.hstack([W21[h1:], W22[h1:, :h1]])                                               # W[b', < b']
    Wbb = W22[h1:, h1:]
    Kinv[:m, :m] += R.T @ R                                                                # Z2a
    Kinv[m:, :m] = Wbb.T @ R                                                               # Z2b
    Kinv[m:, m:] = Wbb.T @ Wbb                                                             # Z2c
    ref = np.linalg.inv(A)
    assert np.abs(np.tril(Kinv) - np.tril(ref)).max() < 1e-12
    # alpha straight from the triangular inverse (trmv_lower_kernel twice)
    y = np.arange(1.0, n1 + n2 + 1.0)
    assert np.abs(W.T @ (W @ y) - ref @ y).max() < 1e-12
