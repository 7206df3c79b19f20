"""CPU: the native host LSAP (amoe_lsap_batched_host) returns exactly scipy's assignment
(scipy.optimize.linear_sum_assignment is what the reference calls, hungarian_matcher.py:79)."""
import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment


def solve(cost, n_tgt, threads=4):
    from automoe_b200 import _ops
    rows, cols, nm = _ops.lsap_batched(torch.as_tensor(cost, dtype=torch.float32), torch.as_tensor(n_tgt), threads)
    return rows.numpy(), cols.numpy(), nm.numpy()


@pytest.mark.parametrize("Q,Nmax,B,seed", [(64, 20, 16, 0), (16, 30, 8, 1), (920, 60, 4, 2), (7, 7, 9, 3), (1, 5, 3, 4), (5, 1, 3, 5)])
def test_lsap_equals_scipy_random(built_lib, Q, Nmax, B, seed):
    rng = np.random.default_rng(seed)
    cost = rng.standard_normal((B, Q, Nmax)).astype(np.float32)
    n_tgt = rng.integers(0, Nmax + 1, size=B).astype(np.int32)
    n_tgt[0] = Nmax
    rows, cols, nm = solve(cost, n_tgt)
    for b in range(B):
        r, c = linear_sum_assignment(cost[b, :, :n_tgt[b]])
        assert nm[b] == len(r) == min(Q, n_tgt[b])
        assert np.array_equal(rows[b, :nm[b]], r) and np.array_equal(cols[b, :nm[b]], c)


def test_lsap_ties_integer_costs(built_lib):
    """Small integer costs create many ties; the tie rule must match scipy's."""
    rng = np.random.default_rng(11)
    for Q, N in [(12, 12), (20, 8), (8, 20), (30, 30)]:
        cost = rng.integers(0, 3, size=(6, Q, N)).astype(np.float32)
        cost[5] = 1.0  # constant matrix -> identity assignment in scipy
        rows, cols, nm = solve(cost, np.full(6, N, np.int32))
        for b in range(6):
            r, c = linear_sum_assignment(cost[b])
            assert np.array_equal(rows[b, :nm[b]], r) and np.array_equal(cols[b, :nm[b]], c)


def test_lsap_empty_and_invalid(built_lib):
    cost = np.zeros((2, 4, 3), np.float32)
    rows, cols, nm = solve(cost, np.array([0, 3], np.int32))
    assert nm.tolist() == [0, 3]
    bad = cost.copy()
    bad[1, 2, 1] = np.nan
    with pytest.raises(ValueError, match="invalid numeric"):
        solve(bad, np.array([3, 3], np.int32))
    bad[1, 2, 1] = -np.inf
    with pytest.raises(ValueError):
        solve(bad, np.array([3, 3], np.int32))
    # scipy accepts +inf entries as long as a finite assignment exists
    ok = np.random.default_rng(0).standard_normal((1, 4, 4)).astype(np.float32)
    ok[0, 0, 0] = np.inf
    rows, cols, nm = solve(ok, np.array([4], np.int32))
    r, c = linear_sum_assignment(ok[0])
    assert np.array_equal(rows[0], r) and np.array_equal(cols[0], c)
