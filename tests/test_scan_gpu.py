"""The library's exclusive prefix sum (csrc/scan.cu), driven directly through nans_debug_scan: every offset table of
the step (cell starts, pair offsets, contact compaction, incidence lists, runs) comes out of it.  Against numpy, from one
element to thousands of tiles, at tile borders and ragged ends, twice in a row (the re-armed epoch)."""
import ctypes as C

import numpy as np
import pytest

from nans_projekat_b200 import _lib

pytestmark = pytest.mark.gpu


def _scan(a):
    a = np.ascontiguousarray(a, np.uint32)
    out = np.zeros_like(a)
    _lib.check(_lib.lib().nans_debug_scan(a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), len(a)))
    return out


@pytest.mark.parametrize("n", [0, 1, 31, 4095, 4096, 4097, 8192, 100_003, 1 << 20, 2_000_001, (1 << 22) - 1, 1 << 22,
                               (1 << 22) + 1, 4096 * 1025, 6_000_011])
def test_exclusive_scan_matches_numpy(n):
    rng = np.random.default_rng(n + 5)
    a = rng.integers(0, 25, n, dtype=np.uint32)
    want = np.zeros(n, np.uint32)
    if n:
        want[1:] = np.cumsum(a[:-1], dtype=np.uint64).astype(np.uint32)
    got = _scan(a)
    assert np.array_equal(got, want), f"n={n}: first difference at {int(np.flatnonzero(got != want)[0])}"


def test_scan_wraps_like_uint32():
    a = np.full(70_000, 0xFFFFFFF0 // 7, np.uint32)
    want = np.zeros_like(a)
    want[1:] = (np.cumsum(a[:-1].astype(np.uint64)) & 0xFFFFFFFF).astype(np.uint32)
    assert np.array_equal(_scan(a), want)
