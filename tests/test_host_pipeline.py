"""Host-side logic of the pipelined host-buffer stepper (no GPU): the slice layout of ShardedHostEnv."""
import numpy as np
import pytest

import ccbs_b200  # noqa: F401  (registers the package under its import name)
from ccbs_b200.dist import shard_range
from ccbs_b200.host_pipeline import slice_bounds


def _covers(bounds, n):
    assert bounds[0][0] == 0 and bounds[-1][1] == n
    for (a, b), (c, d) in zip(bounds, bounds[1:]):
        assert b == c and b > a and d > c
    return True


def test_equal_slices_are_the_rank_shards():
    for n, k in ((8192, 4), (10, 3), (5, 5), (4097, 8)):
        b = slice_bounds(n, k)
        assert b == [shard_range(n, r, k) for r in range(k)]
        assert _covers(b, n)


def test_weighted_slices_cover_the_batch_in_proportion():
    b = slice_bounds(8192, 5, [30, 30, 25, 12, 3])
    assert _covers(b, 8192)
    sizes = np.array([hi - lo for lo, hi in b])
    assert np.all(np.abs(sizes / 8192 - np.array([0.30, 0.30, 0.25, 0.12, 0.03])) < 1e-3)
    assert np.all(np.diff(sizes) <= 0)          # tapering: the last slice (whose kernels run after the link has gone idle) is the smallest


def test_every_slice_keeps_an_env():
    b = slice_bounds(5, 5, [10, 1, 1, 1, 1e-3])
    assert b == [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5)]
    b = slice_bounds(7, 3, [1e-6, 1, 1e-6])
    assert _covers(b, 7) and min(hi - lo for lo, hi in b) >= 1


def test_weights_must_be_positive_and_match_the_slice_count():
    with pytest.raises(ValueError):
        slice_bounds(100, 3, [1, 0, 1])
    assert slice_bounds(100, 4, [1, 2]) == [shard_range(100, r, 4) for r in range(4)]   # wrong length: ignored


def test_order_preserving_integer_image_of_a_float():
    """decode_select takes the warp maximum of float32 scores with one integer REDUX (k_decode.cu: ci ^= (ci >> 31) & 0x7FFFFFFF before
    and after): the image must order like the floats (NaN never reaches it; -0.0 sorts below +0.0, which fmaxf may return either way)."""
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(4096).astype(np.float32) * np.float32(10.0) ** rng.integers(-30, 30, 4096).astype(np.float32),
                        np.array([0.0, -0.0, np.inf, -np.inf, 1e-45, -1e-45, 3.4e38, -3.4e38], dtype=np.float32)])
    i = x.view(np.int32).copy()
    img = i ^ ((i >> 31) & np.int32(0x7FFFFFFF))
    order = np.argsort(img, kind="stable")
    xs = x[order]
    assert np.all(np.diff(xs.astype(np.float64)) >= 0)
    back = img ^ ((img >> 31) & np.int32(0x7FFFFFFF))
    assert np.array_equal(back, i)                                  # the map is its own inverse
    assert x[np.argmax(img)] == x.max() and x[np.argmin(img)] == x.min()
