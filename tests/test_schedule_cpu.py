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
