"""Fisher exact test.  Pin: exact integer / rational evaluation of the published kt_fisher_exact procedure
(oracle/fisher_exact_rational.py -> tests/golden/fisher_exact.tsv.gz, 10 160 tables incl. exact ties and margins up to
10 000).  Against it, at 1e-12 (n <= 100) / 1e-10 (n <= 1200) / 1e-9 relative: the double-precision restatement oracle/fisher.py, the
kt_fisher_exact the reference binary links (oracle/hts_shim), and the CUDA kernel.  scipy stays as a second opinion."""
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


def _golden_fisher():
    import gzip
    rows = []
    with gzip.open(os.path.join(ROOT, "tests", "golden", "fisher_exact.tsv.gz"), "rt") as f:
        for l in f:
            t = l.split("\t")
            rows.append((tuple(int(x) for x in t[:4]), tuple(float(x) for x in t[4:7])))
    return rows


def _tol(tab):
    # the double-precision evaluations go through exp(sum of lgamma): their relative error grows with |log p|, i.e. with
    # the table's size (~ n * 1e-16 * a few); the exact vectors show it directly
    n = sum(tab)
    return 1e-12 if n <= 100 else 1e-10 if n <= 1200 else 1e-9


def test_golden_fisher_file_is_the_exact_procedure():
    """The committed vectors are what oracle/fisher_exact_rational.py computes (spot-check: every 37th table)."""
    import fisher_exact_rational as ex
    rows = _golden_fisher()
    assert len(rows) >= 10000 and [r[0] for r in rows] == ex.tables()
    for tab, want in rows[::37]:
        got = tuple(float(x) for x in ex.exact_fisher(*tab))
        assert got == want, tab


def test_oracle_fisher_matches_exact_rational():
    for tab, want in _golden_fisher():
        got = oracle_fisher.kt_fisher_exact(*tab)
        for x, y in zip(want, got):
            assert y == pytest.approx(x, rel=_tol(tab), abs=1e-300), (tab, want, got)


def test_reference_side_shim_fisher_matches_exact_rational(tmp_path):
    """oracle/hts_shim's kt_fisher_exact — the function the compiled reference (oracle/_ref) calls, i.e. the source of the
    Fisher columns in every golden TSV — against the exact vectors."""
    import ctypes as C
    import subprocess
    so = str(tmp_path / "libshim.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-std=gnu++11", "-O2", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "oracle", "hts_shim"),
                    os.path.join(ROOT, "oracle", "hts_shim", "hts_shim.cpp"), "-o", so, "-lz"], check=True)
    L = C.CDLL(so)
    L.kt_fisher_exact.restype = C.c_double
    L.kt_fisher_exact.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_double)] * 3
    for tab, want in _golden_fisher():
        l, r, t = C.c_double(), C.c_double(), C.c_double()
        L.kt_fisher_exact(*tab, C.byref(l), C.byref(r), C.byref(t))
        for x, y in zip(want, (l.value, r.value, t.value)):
            assert y == pytest.approx(x, rel=_tol(tab), abs=1e-300), (tab, want)


@pytest.mark.gpu
def test_cuda_fisher_matches_exact_rational(built):
    import rabbitvar_b200 as rv
    rows = _golden_fisher()
    ctx = rv.Context(0)
    got = ctx.fisher_exact(np.array([r[0] for r in rows], dtype=np.int32))
    ctx.close()
    for (tab, want), g in zip(rows, got):
        for x, y in zip(want, g):
            assert y == pytest.approx(x, rel=_tol(tab), abs=1e-300), (tab, want, tuple(g))
