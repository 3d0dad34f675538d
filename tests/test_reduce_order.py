"""The summation order of one partial record, restated in numpy (no GPU): csrc/reduce.cuh reduces the per-lane sums of a warp
with a PACKED butterfly -- at lane distance o the two lanes of a pair split the values still to be reduced, so 33 floats need
17 + 9 + 5 + 3 + 2 shuffles instead of 5 x 33.  What the kernels rely on, and what this test pins down on the algorithm itself:
  * every value's complete sum ends up on at least one lane, under the row index the code derives for it;
  * that sum is the plain butterfly tree (L, L^16), (., .^8), (., .^4), (., .^2), (., .^1) in fp32 -- a pure function of the 32
    inputs, identical on every lane that holds it (fp addition is commutative), so the record is bit-reproducible;
  * the shuffle count is what DESIGN.md states.
"""
import numpy as np
import pytest


def packed_step(a, ia, o):
    """a: (32, N) float32 per-lane values, ia: (32, N) row ids.  Returns ((32, ceil(N/2)) values, ids, shuffles)."""
    lanes, n = a.shape
    upper = (np.arange(lanes) & o) != 0
    partner = np.arange(lanes) ^ o
    nb = (n + 1) // 2
    b = np.zeros((lanes, nb), a.dtype)
    ib = np.zeros((lanes, nb), np.int64)
    shuffles = 0
    for i in range(n // 2):
        keep = np.where(upper, a[:, 2 * i + 1], a[:, 2 * i])
        give = np.where(upper, a[:, 2 * i], a[:, 2 * i + 1])
        b[:, i] = keep + give[partner]  # __shfl_xor_sync(give, o)
        ib[:, i] = np.where(upper, ia[:, 2 * i + 1], ia[:, 2 * i])
        shuffles += 1
    if n & 1:
        b[:, nb - 1] = a[:, n - 1] + a[partner, n - 1]
        ib[:, nb - 1] = ia[:, n - 1]
        shuffles += 1
    return b, ib, shuffles


def butterfly(col):
    """plain butterfly all-reduce of one value over 32 lanes (fp32), the tree the packed version must reproduce"""
    v = col.copy()
    for o in (16, 8, 4, 2, 1):
        v = v + v[np.arange(32) ^ o]
    return v


@pytest.mark.parametrize("n,dtype,expect_shuffles", [(33, np.float32, 36), (46, np.float32, 23 + 12 + 6 + 3 + 2), (5, np.float64, 8)])
def test_packed_butterfly_is_the_plain_butterfly_tree(n, dtype, expect_shuffles):
    rng = np.random.default_rng(n)
    vals = (rng.standard_normal((32, n)) * 10.0 ** rng.integers(-3, 4, (32, n))).astype(dtype)
    a, ia = vals.copy(), np.tile(np.arange(n), (32, 1))
    total = 0
    for o in (16, 8, 4, 2, 1):
        a, ia, s = packed_step(a, ia, o)
        total += s
    assert total == expect_shuffles
    want = np.stack([butterfly(vals[:, r]) for r in range(n)], axis=1)  # (32, n): identical on all lanes up to commutativity
    seen = set()
    for lane in range(32):
        for j in range(a.shape[1]):
            r = int(ia[lane, j])
            seen.add(r)
            assert a[lane, j] == want[lane, r], (lane, j, r)  # bit-exact: same tree
            assert a[lane, j] == want[0, r]                   # and the same bits on every lane that holds the row
    assert seen == set(range(n)), "every row's sum must be held by some lane"
