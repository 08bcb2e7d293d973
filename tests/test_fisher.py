"""Fisher exact test: the oracle restatement is pinned against scipy; the CUDA kernel against the oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fisher as oracle_fisher  # noqa: E402


def _tables(seed=3, n=400):
    rng = np.random.default_rng(seed)
    t = [(0, 0, 0, 0), (1, 0, 0, 1), (5, 0, 0, 5), (10, 10, 10, 10), (50, 48, 3, 0), (30, 26, 30, 26),
         (2000, 1900, 12, 30), (4000, 4100, 25, 20), (0, 7, 9, 0), (100, 0, 100, 0)]
    for _ in range(n):
        depth = int(rng.choice([20, 100, 500, 5000]))
        a, b = rng.integers(0, depth, 2)
        c, d = rng.integers(0, max(2, depth // 10), 2)
        t.append((int(a), int(b), int(c), int(d)))
    return t


def test_oracle_fisher_matches_scipy():
    from scipy.stats import fisher_exact
    for a, b, c, d in _tables():
        two = oracle_fisher.kt_fisher_exact(a, b, c, d)[2]
        want = fisher_exact([[a, b], [c, d]])[1]
        # kfunc's two-sided p differs from scipy's only through its 1e-8 tie window
        assert two == pytest.approx(want, rel=1e-6, abs=1e-300), (a, b, c, d)


@pytest.mark.gpu
def test_cuda_fisher_matches_oracle(built):
    import rabbitvar_b200 as rv
    ctx = rv.Context(0)
    t = _tables()
    got = ctx.fisher_exact(np.array(t, dtype=np.int32))
    for (a, b, c, d), g in zip(t, got):
        want = oracle_fisher.kt_fisher_exact(a, b, c, d)
        for x, y in zip(want, g):
            assert y == pytest.approx(x, rel=1e-9, abs=1e-300), (a, b, c, d)
    ctx.close()
