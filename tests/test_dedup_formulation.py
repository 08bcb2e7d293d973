"""The -t filter: the reference's running set (oracle/dedup.py::running_set_filter, recordPreprocessor.cpp:153-176)
and the order-free scan the CUDA path uses agree on coordinate-sorted record streams."""
import os
import random
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import dedup  # noqa: E402


def _stream(rng, n):
    recs = []
    pos = 100
    cigars = [(150 << 4,), ((100 << 4), (50 << 4) | 4), ((10 << 4) | 4, (140 << 4)), ((75 << 4), (2 << 4) | 1, (73 << 4))]
    for _ in range(n):
        pos += rng.choice([0, 0, 0, 1, 2])
        paired = rng.random() < 0.7
        flag = (1 if paired else 0) | (4 if rng.random() < 0.3 else 0) | (16 if rng.random() < 0.5 else 0)
        mtid = rng.choice([0, 0, 1, 2, -1])
        mpos = rng.choice([0, 0, 5, 9, 10, 11, pos + 200])
        recs.append(dict(pos=pos, mpos=mpos, flag=flag, tid=0, mtid=mtid, cigar=rng.choice(cigars)))
    return recs


def test_scan_back_equals_running_set_on_sorted_streams():
    rng = random.Random(20261017)
    dropped = dropped_cigar_key = 0
    for _ in range(300):
        recs = _stream(rng, rng.randint(0, 60))
        want = dedup.running_set_filter(recs)
        got = dedup.scan_back_filter(recs)
        assert got == want, (recs, want, got)
        dropped += want.count(False)
        dropped_cigar_key += sum(1 for r, k in zip(recs, want) if not k and dedup.key_kind(r) == 2)
    assert dropped > 200 and dropped_cigar_key > 10  # the streams do exercise both key kinds


def test_empty_and_single():
    assert dedup.scan_back_filter([]) == dedup.running_set_filter([]) == []
    r = [dict(pos=5, mpos=0, flag=0, tid=0, mtid=-1, cigar=(150 << 4,))]
    assert dedup.scan_back_filter(r) == dedup.running_set_filter(r) == [True]
