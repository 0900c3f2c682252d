"""GPU parity: per-block model selection (first minimum = lowest id on ties; KF.py:821-864, V22.py:2233-2252 / 2350-2369)."""
import numpy as np
import pytest

import datasets
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_select_first_minimum_with_ties_and_skipped_candidates():
    import gpu_util as G
    rng = np.random.default_rng(11)
    big = np.iinfo(np.int64).max
    for nb, nc in ((1, 1), (1, 4), (7, 10), (1000, 11), (4096, 4)):
        s = rng.integers(0, 6, size=(nb, nc)).astype(np.int64)          # few values: many ties
        s[rng.random((nb, nc)) < 0.2] = big                                # candidates the reference skipped
        s[:, 0] = np.minimum(s[:, 0], 1 << 40)                             # raw is always there
        mids, best = G.ctx().select_blocks(s)
        assert (mids == np.argmin(s, axis=1)).all()
        assert (best == s.min(axis=1)).all()
    mids, best = G.ctx().select_blocks(np.zeros((0, 4), dtype=np.int64))
    assert len(mids) == 0 and len(best) == 0


def test_select_agrees_with_the_oracle_selection_on_real_sizes():
    """The sizes the oracle reports for every candidate of a block, pushed through the device selection, give the oracle's method."""
    import gpu_util as G
    rows, want = [], []
    for name, d in sorted(datasets.small_cases().items()):
        if not d:
            continue
        for prof in (O.PROFILE_KOLM, O.PROFILE_KOLR):
            mid, _, sizes = O.encode_block(prof, d)
            if prof == O.PROFILE_KOLM:
                rows.append([s if s is not None else np.iinfo(np.int64).max for s in sizes] + [np.iinfo(np.int64).max] * 6)
                want.append(mid)
            else:
                rows.append([s if s is not None else np.iinfo(np.int64).max for s in sizes])
                want.append(mid)
    mids, _ = G.ctx().select_blocks(np.array(rows, dtype=np.int64))
    assert mids.tolist() == want
