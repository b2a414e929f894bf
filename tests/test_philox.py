"""Host Philox restatement (ccbs_b200/philox.py) against the scalar one the GPU lockstep tests use to predict the device's draws,
and the sub-sampling rule that stands in for np.random.choice in __balance_action_space_by_outcome (compressed:553-567)."""
import numpy as np
import pytest

from ccbs_b200.philox import philox4x32_10, row_identity, subset_keep
from tests.test_gpu_lockstep import _philox_words


def test_vectorised_philox_matches_scalar():
    rng = np.random.default_rng(0)
    for _ in range(20):
        seed, env = int(rng.integers(0, 2 ** 63)), int(rng.integers(0, 2 ** 40))
        steps, streams = rng.integers(0, 2 ** 32, size=7), rng.integers(0, 2 ** 32, size=7)
        w = philox4x32_10(seed, env, steps, streams)
        assert w.shape == (7, 4) and w.dtype == np.uint32
        for i in range(7):
            assert w[i].tolist() == _philox_words(seed, env, int(steps[i]), int(streams[i]))


def test_subset_keep_is_a_uniform_ordered_subset():
    ident = row_identity([3] * 40, [5] * 40, [2] * 40, np.arange(40))
    seen = np.zeros(40)
    for call in range(400):
        keep = subset_keep(9, 1, call, ident, 10)
        assert len(keep) == 10 and np.all(np.diff(keep) > 0) and keep.min() >= 0 and keep.max() < 40
        seen[keep] += 1
    assert seen.min() > 60 and seen.max() < 140                 # each row kept ~100 times out of 400
    # deterministic in (seed, env, call); different calls / envs draw different subsets
    assert np.array_equal(subset_keep(9, 1, 7, ident, 10), subset_keep(9, 1, 7, ident, 10))
    assert not np.array_equal(subset_keep(9, 1, 7, ident, 10), subset_keep(9, 2, 7, ident, 10))
    # rows with the same identity share the key and keep their table order: the first of them wins
    dup = row_identity([1, 1, 1, 1], [2, 2, 2, 2], [0, 0, 0, 0], [6, 6, 7, 7])
    for call in range(10):
        keep = subset_keep(3, 0, call, dup, 2).tolist()
        assert keep in ([0, 1], [2, 3], [0, 2])
    with pytest.raises(ValueError):
        row_identity([128], [0], [0], [0])
