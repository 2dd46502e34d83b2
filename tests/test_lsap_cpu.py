"""CPU: gwd_lsap_batch (the native, multi-threaded assignment solver of the matcher) returns index-for-index what
scipy.optimize.linear_sum_assignment returns -- random costs, heavily TIED costs (small integers, duplicated rows /
columns: equal-cost optima, where only the same algorithm with the same tie rules agrees), all aspect ratios, empty
and single-entry problems, and the matcher's own shape (100 queries x 12..47 targets)."""
import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment

import gwdepth_b200  # noqa: F401
from gwdepth_b200 import capi, ops

pytestmark = pytest.mark.skipif(not __import__("os").path.exists(capi.LIB_PATH), reason="libgwd_b200.so not built")


def solve_all(mats, Q, threads=0):
    flat = np.concatenate([m.reshape(-1) for m in mats]) if mats else np.zeros(0, np.float32)
    offs, o = [], 0
    for m in mats:
        offs.append(o)
        o += m.size
    return ops.lsap_batch(flat, offs, [m.shape[1] for m in mats], Q, n_threads=threads)


def check(mats, Q, threads=0):
    got = solve_all(mats, Q, threads)
    for m, (qi, ti) in zip(mats, got):
        ri, ci = linear_sum_assignment(m)
        assert np.array_equal(qi, ri) and np.array_equal(ti, ci), (m.shape, qi, ri, ti, ci)


@pytest.mark.parametrize("Q", [100, 7, 1])
def test_random_costs_all_aspect_ratios(Q):
    rng = np.random.default_rng(Q)
    mats = [rng.random((Q, T)).astype(np.float32) * 5 - rng.random((Q, 1)).astype(np.float32) for T in list(range(1, 60)) + [100, 130]]
    check(mats, Q)
    check(mats[:5], Q, threads=1)


@pytest.mark.parametrize("Q", [100, 12])
def test_tied_costs_pick_the_same_optimum_as_scipy(Q):
    rng = np.random.default_rng(100 + Q)
    mats = []
    for T in (1, 2, 5, 12, 17, 47, 100, 120):
        mats.append(rng.integers(0, 3, (Q, T)).astype(np.float32))              # many equal entries
        mats.append(np.zeros((Q, T), np.float32))                               # everything tied
        m = rng.random((Q, T)).astype(np.float32)
        m[Q // 2:] = m[: Q - Q // 2]                                            # duplicated queries (collapsed predictions)
        mats.append(m)
        m = rng.random((Q, T)).astype(np.float32)
        m[:, T // 2:] = m[:, : T - T // 2]                                      # duplicated targets
        mats.append(m)
    check(mats, Q)


def test_matcher_shaped_batch_and_degenerate_problems():
    rng = np.random.default_rng(7)
    mats = [(5 * rng.random((100, 12 + 5 * (b % 8))) - rng.random((100, 1))).astype(np.float32) for b in range(96)]
    check(mats, 100)
    got = solve_all([np.zeros((100, 0), np.float32), rng.random((100, 3)).astype(np.float32)], 100)
    assert len(got[0][0]) == 0 and len(got[1][0]) == 3
    with pytest.raises(capi.GwdError):
        solve_all([np.full((4, 3), np.nan, np.float32)], 4)
