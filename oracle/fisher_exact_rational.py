"""Exact evaluation of the published kt_fisher_exact procedure — TEST INFRASTRUCTURE (oracle).

htslib's kfunc.c (un-vendored, unpinned dependency of the reference: install.sh:40; call sites
src/modes/simpleMode.cpp:96-108, src/modes/somaticMode.cpp:130-149) defines the two-sided p of a 2x2 table as the
sum of the hypergeometric probabilities of the two tails, each tail walked inwards from its end while
p < 0.99999999 q (q = probability of the observed table), plus the boundary term when p < 1.00000001 q; left / right
are the one-sided sums, the far one replaced by 1 - near + q.

Here every probability is an exact integer ratio (numerators C(n1_, k) C(n - n1_, n_1 - k) over the common denominator
C(n, n_1), built by the exact integer recurrence), every comparison of the procedure is decided in integer arithmetic
and every sum is exact; only the final results are rounded to double.  This pins the double-precision restatements
(oracle/fisher.py, the shim the reference binary links: oracle/hts_shim/hts_shim.cpp, and the CUDA kernel
rv_fisher_kernel / fisher_exact in csrc/kernels/rv_score.cuh) independently of any lgamma: none of them shares
code or arithmetic with this file.
"""
from fractions import Fraction
from math import comb


def exact_fisher(n11, n12, n21, n22):
    """(left, right, two) of kt_fisher_exact as exact Fractions."""
    n1_, n_1, n = n11 + n12, n11 + n21, n11 + n12 + n21 + n22
    mx = min(n_1, n1_)
    mn = max(n1_ + n_1 - n, 0)
    if mn == mx:
        return Fraction(1), Fraction(1), Fraction(1)
    # numerators of P(k), k = mn..mx, by the exact recurrence num(k+1) = num(k) (n1_-k)(n_1-k) / ((k+1)(k+1+n-n1_-n_1))
    num = [comb(n1_, mn) * comb(n - n1_, n_1 - mn)]
    for k in range(mn, mx):
        a = num[-1] * (n1_ - k) * (n_1 - k)
        b = (k + 1) * (k + 1 + n - n1_ - n_1)
        assert a % b == 0
        num.append(a // b)
    den = comb(n, n_1)
    assert sum(num) == den
    q = num[n11 - mn]
    lo, hi = 99999999, 100000001  # p < 0.99999999 q  <=>  1e8 p < 99999999 q ; p < 1.00000001 q  <=>  1e8 p < 100000001 q
    P = lambda k: num[k - mn]
    # left tail
    left = 0
    i = mn + 1
    p = P(mn)
    while 100000000 * p < lo * q and i <= mx:
        left += p
        p = P(i)
        i += 1
    i -= 1
    if 100000000 * p < hi * q:
        left += p
    else:
        i -= 1
    # right tail
    right = 0
    j = mx - 1
    p = P(mx)
    while 100000000 * p < lo * q and j >= 0:
        right += p
        p = P(j) if j >= mn else 0
        j -= 1
    j += 1
    if 100000000 * p < hi * q:
        right += p
    else:
        j += 1
    two = Fraction(left + right, den)
    if two > 1:
        two = Fraction(1)
    L, R, Q = Fraction(left, den), Fraction(right, den), Fraction(q, den)
    if abs(i - n11) < abs(j - n11):
        R = 1 - L + Q
    else:
        L = 1 - R + Q
    return L, R, two


def tables(seed=20261017, n_small=10000, n_large=160):
    """The pinned table set: exhaustive tiny tables, ties (symmetric margins), random tables at the depths of the five
    configs, margins up to 10 000."""
    import random
    rnd = random.Random(seed)
    t = [(a, b, c, d) for a in range(5) for b in range(5) for c in range(5) for d in range(5)]
    t += [(k, k, k, k) for k in (1, 2, 7, 30, 100, 400)] + [(k, m, m, k) for k in (3, 10, 50) for m in (1, 5, 60)]
    t += [(10, 10, 10, 10), (50, 48, 3, 0), (30, 26, 30, 26), (2000, 1900, 12, 30), (4000, 4100, 25, 20), (0, 7, 9, 0),
          (100, 0, 100, 0), (5000, 5000, 0, 7), (1, 9999, 9999, 1)]
    while len(t) < n_small:
        depth = rnd.choice([8, 20, 60, 100, 200, 300])
        a, b = rnd.randrange(depth + 1), rnd.randrange(depth + 1)
        c, d = rnd.randrange(max(2, depth // 4)), rnd.randrange(max(2, depth // 4))
        t.append((a, b, c, d))
    for _ in range(n_large):
        depth = rnd.choice([1000, 2500, 5000])
        a, b = rnd.randrange(depth + 1), rnd.randrange(depth + 1)
        c, d = rnd.randrange(depth // 20 + 2), rnd.randrange(depth // 20 + 2)
        t.append((a, b, c, d))
    return t


if __name__ == "__main__":  # writes the committed golden vector file (tests/golden/fisher_exact.tsv.gz)
    import gzip
    import os
    import sys
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "fisher_exact.tsv.gz")
    with gzip.open(out, "wt") as f:
        for a, b, c, d in tables():
            L, R, T = exact_fisher(a, b, c, d)
            f.write(f"{a}\t{b}\t{c}\t{d}\t{float(L)!r}\t{float(R)!r}\t{float(T)!r}\n")
    print("wrote", out)
