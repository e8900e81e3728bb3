"""Host logic of the INT8 split path (csrc/api_ozaki.cu, dsmgp_host_split_plan): which diagonal ranges an expert is cut into and how
much of its factorisation + inverse work the block products carry.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

from deepstructuredmixtures_b200 import _native as nat


def plan(n, depth=1, min_nb=8):
    lib = nat.lib()
    nb = (((n + 63) // 64) * 64 + 127) // 128
    ro = np.zeros(nb, dtype=np.int32)
    nr = C.c_int32(0)
    share = C.c_double(0.0)
    rc = lib.dsmgp_host_split_plan(n, depth, min_nb, ro.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(nr), C.byref(share))
    assert rc == 0
    return ro, nr.value, share.value


def test_small_experts_are_not_split():
    for n in (10, 200, 640, 896):                # fewer than 8 block rows
        ro, nr, share = plan(n)
        assert nr == 1 and np.all(ro == 0) and share == pytest.approx(0.0, abs=1e-15)


@pytest.mark.parametrize("n", [1024, 1030, 1664, 2500, 5008, 8544])
def test_one_level_split_at_the_middle_block_row(n):
    ro, nr, share = plan(n)
    nb = len(ro)
    assert nr == 2
    mid = nb // 2
    assert np.all(ro[:mid] == 0) and np.all(ro[mid:] == 1)
    r1 = min(n, mid * 128); r2 = n - r1
    assert share == pytest.approx(1.0 - (r1 / n) ** 3 - (r2 / n) ** 3, rel=1e-12)
    assert 0.70 <= share <= 0.75 + 1e-12         # 3/4 of the flops for an even split


def test_two_levels_and_the_minimum_size():
    ro, nr, share = plan(5008, depth=2, min_nb=8)     # 40 block rows: 20 | 20, then 10 | 10 each
    assert nr == 4 and [int((ro == k).sum()) for k in range(4)] == [10, 10, 10, 10]
    assert share > 0.93
    ro, nr, _ = plan(2500, depth=2, min_nb=8)         # 20 block rows: 10 | 10, and 10 >= 8 splits once more
    assert nr == 4
    ro, nr, _ = plan(1664, depth=2, min_nb=8)         # 13 block rows: 6 | 7, neither half has 8 rows
    assert nr == 2 and int((ro == 0).sum()) == 6
    ro, nr, _ = plan(1664, depth=2, min_nb=14)
    assert nr == 1


def test_ranges_are_contiguous_and_ordered():
    for n in (1500, 3333, 7000):
        for depth in (1, 2, 3):
            ro, nr, _ = plan(n, depth=depth, min_nb=4)
            assert ro[0] == 0 and np.all(np.diff(ro) >= 0) and np.all(np.diff(ro) <= 1) and ro[-1] == nr - 1
