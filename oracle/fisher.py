"""CPU restatement of kt_fisher_exact — TEST INFRASTRUCTURE (oracle).

The reference calls htslib's kt_fisher_exact (call sites src/modes/simpleMode.cpp:98,
src/modes/somaticMode.cpp:132).  htslib is an un-vendored, unpinned dependency (install.sh:40 clones
samtools/htslib HEAD), so the arithmetic is restated from the published algorithm of kfunc.c:
    lbinom(n,k)  = lgamma(n+1) - lgamma(k+1) - lgamma(n-k+1)
    hypergeo     = exp(lbinom(n1_,n11) + lbinom(n-n1_,n_1-n11) - lbinom(n,n_1))
incremental ratio updates between exact re-evaluations at every 11th step, tails summed while
p < 0.99999999 q, boundary term added when p < 1.00000001 q, two = min(1, left + right).
The reference's own tests do not pin these values (it has none); they are pinned against the exact integer
evaluation of the same procedure (oracle/fisher_exact_rational.py, tests/golden/fisher_exact.tsv.gz) at 1e-12 and
cross-checked with scipy.stats.fisher_exact in tests/test_fisher.py.
"""
from math import exp, lgamma


def _lbinom(n, k):
    if k == 0 or n == k:
        return 0.0
    return lgamma(n + 1) - lgamma(k + 1) - lgamma(n - k + 1)


def _hypergeo(n11, n1_, n_1, n):
    return exp(_lbinom(n1_, n11) + _lbinom(n - n1_, n_1 - n11) - _lbinom(n, n_1))


class _Acc:
    __slots__ = ("n11", "n1_", "n_1", "n", "p")


def _hypergeo_acc(n11, n1_, n_1, n, a):
    if n1_ or n_1 or n:
        a.n11, a.n1_, a.n_1, a.n = n11, n1_, n_1, n
    else:
        if n11 % 11 and n11 + a.n - a.n1_ - a.n_1:
            if n11 == a.n11 + 1:
                a.p *= (a.n1_ - a.n11) / n11 * (a.n_1 - a.n11) / (n11 + a.n - a.n1_ - a.n_1)
                a.n11 = n11
                return a.p
            if n11 == a.n11 - 1:
                a.p *= a.n11 / (a.n1_ - n11) * (a.n11 + a.n - a.n1_ - a.n_1) / (a.n_1 - n11)
                a.n11 = n11
                return a.p
        a.n11 = n11
    a.p = _hypergeo(a.n11, a.n1_, a.n_1, a.n)
    return a.p


def kt_fisher_exact(n11, n12, n21, n22):
    """Returns (left, right, two_sided)."""
    a = _Acc()
    n1_, n_1, n = n11 + n12, n11 + n21, n11 + n12 + n21 + n22
    mx = min(n_1, n1_)
    mn = max(n1_ + n_1 - n, 0)
    if mn == mx:
        return 1.0, 1.0, 1.0
    q = _hypergeo_acc(n11, n1_, n_1, n, a)
    p = _hypergeo_acc(mn, 0, 0, 0, a)
    left = 0.0
    i = mn + 1
    while p < 0.99999999 * q and i <= mx:
        left += p
        p = _hypergeo_acc(i, 0, 0, 0, a)
        i += 1
    i -= 1
    if p < 1.00000001 * q:
        left += p
    else:
        i -= 1
    p = _hypergeo_acc(mx, 0, 0, 0, a)
    right = 0.0
    j = mx - 1
    while p < 0.99999999 * q and j >= 0:
        right += p
        p = _hypergeo_acc(j, 0, 0, 0, a)
        j -= 1
    j += 1
    if p < 1.00000001 * q:
        right += p
    else:
        j += 1
    two = min(1.0, left + right)
    if abs(i - n11) < abs(j - n11):
        right = 1.0 - left + q
    else:
        left = 1.0 - right + q
    return left, right, two
