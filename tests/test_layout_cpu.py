"""Host logic of the fast mode that needs no GPU: the phase-separated internal layout (urnn_layout_index).

The layout is what turns AvgPool2 (utils.py:92-94) and ConvTranspose2d(k2,s2) (utils.py:95-100) into pure channel
operations: the four 2x2 children of a coarse pixel sit at the coarse pixel's own position inside four consecutive block
groups of the finer map."""
import ctypes

import numpy as np
import pytest

from urnn_b200 import _capi


def index_map(lib, H, W, level):
    n = ctypes.c_int64(0)
    h, w = H >> level, W >> level
    idx = np.array([[lib.urnn_layout_index(H, W, level, y, x, ctypes.byref(n)) for x in range(w)] for y in range(h)])
    return idx, n.value


@pytest.mark.parametrize("H,W", [(8, 12), (20, 36), (500, 500)])
def test_layout_is_a_padded_bijection_and_nests(H, W):
    lib = _capi.load()
    if H * W > 10000:                      # full-size grid: spot checks only
        n = ctypes.c_int64(0)
        assert lib.urnn_layout_index(H, W, 0, H - 1, W - 1, ctypes.byref(n)) < n.value
        assert n.value == 16 * ((H // 4) * (W // 4) + 127) // 128 * 128 or n.value % 128 == 0
        assert n.value == 16 * (((H // 4) * (W // 4) + 127) // 128 * 128)
        return
    maps, ntot = [], []
    for level in range(3):
        idx, n = index_map(lib, H, W, level)
        maps.append(idx); ntot.append(n)
        assert idx.min() >= 0 and idx.max() < n
        assert len(np.unique(idx)) == idx.size                 # no two pixels share a slot
    n4p = ntot[2]
    assert n4p % 128 == 0 and n4p >= (H // 4) * (W // 4)
    assert ntot[1] == 4 * n4p and ntot[0] == 16 * n4p          # whole tiles per block: tiles never straddle blocks
    for level in (0, 1):                                       # child (2y+dy, 2x+dx) = phase (dy*2+dx) block + parent position
        fine, coarse = maps[level], maps[level + 1]
        for dy in range(2):
            for dx in range(2):
                np.testing.assert_array_equal(fine[dy::2, dx::2], (dy * 2 + dx) * ntot[level + 1] + coarse)


def test_layout_rejects_bad_requests():
    lib = _capi.load()
    assert lib.urnn_layout_index(10, 12, 0, 0, 0, None) == -1      # H not a multiple of 4
    assert lib.urnn_layout_index(8, 12, 3, 0, 0, None) == -1
    assert lib.urnn_layout_index(8, 12, 1, 4, 0, None) == -1       # row outside the half-resolution map
