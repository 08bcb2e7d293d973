"""Host decode path: BGZF/BAM/BAI/FAI reader against the generator's own bookkeeping."""
import os

import numpy as np

import cases
import rabbitvar_b200 as rv


def test_bam_roundtrip_counts(built):
    d = cases.generate("c1_k1")
    meta = dict(l.split("\t") for l in open(os.path.join(d, "meta.txt")).read().splitlines())
    b = rv.HostBatch(os.path.join(d, "S.bam"), "chrS1", 1, int(meta["len"]))
    assert b.n_reads == int(meta["reads"])
    assert b.chr_len == int(meta["len"])
    reads = b.reads_numpy().view(np.dtype([("pos", "<i4"), ("mpos", "<i4"), ("off", "<u4"), ("l_seq", "<i4"),
                                          ("flag", "<u2"), ("n_cigar", "<u2"), ("nm", "<i2"), ("mapq", "u1"),
                                          ("same", "u1"), ("end", "<i4"), ("rsv", "<i4")]))
    assert (np.diff(reads["pos"]) >= 0).all()          # coordinate sorted
    assert (reads["l_seq"] == 150).all()
    assert (reads["nm"] >= 0).all()                     # NM always present
    assert (reads["end"] >= reads["pos"]).all()


def test_region_query_matches_linear_scan(built):
    d = cases.generate("c1_k1")
    whole = rv.HostBatch(os.path.join(d, "S.bam"), "chrS1", 1, 30000)
    part = rv.HostBatch(os.path.join(d, "S.bam"), "chrS1", 5001, 6000)
    dt = np.dtype([("pos", "<i4"), ("mpos", "<i4"), ("off", "<u4"), ("l_seq", "<i4"), ("flag", "<u2"),
                   ("n_cigar", "<u2"), ("nm", "<i2"), ("mapq", "u1"), ("same", "u1"), ("end", "<i4"), ("rsv", "<i4")])
    w = whole.reads_numpy().view(dt)
    p = part.reads_numpy().view(dt)
    # htslib iterator semantics: pos0 < end && endpos > beg0
    sel = w[(w["pos"] - 1 < 6000) & (w["end"] > 5000)]
    assert len(sel) == len(p)
    assert (sel["pos"] == p["pos"]).all() and (sel["flag"] == p["flag"]).all()


def test_fetch_ref_matches_fasta(built):
    d = cases.generate("c1_k1")
    fa = "".join(open(os.path.join(d, "ref.fa")).read().splitlines()[1:])
    got = rv.fetch_ref(os.path.join(d, "ref.fa"), "chrS1", 1234, 2345).decode()
    assert got == fa[1233:2345]


def test_fixed_point_formatter_matches_libc(tmp_path):
    """csrc/host/fmt.hpp (the %f replacement of the TSV writers) against std::to_string on 3 M values incl. ties."""
    import subprocess
    exe = str(tmp_path / "fmt_test")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(cases.ROOT, "tests", "tools", "fmt_test.cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-500:]


def test_block_decoder_and_crc_match_zlib(built):
    """The loader's own raw-DEFLATE decoder and CRC-32 (csrc/io/fast_inflate.hpp) against zlib: streams of every block
    type (stored, fixed, dynamic), sizes from empty to several blocks, literal- and match-heavy payloads, and every
    BGZF block of a synthetic BAM.  Truncated input and a too-small output buffer must be refused, not overrun."""
    import random
    import struct
    import zlib
    import rabbitvar_b200 as rv
    rnd = random.Random(5)
    payloads = [b"", b"A", bytes(rnd.getrandbits(8) for _ in range(5000)), bytes(rnd.choice(b"ACGT") for _ in range(70000)),
                bytes([37]) * 66000, b"".join(bytes([rnd.choice((37, 37, 37, 30, 25, 12))]) for _ in range(40000)),
                bytes((i * 7 + (i >> 5)) & 255 for i in range(200000))]
    n = 0
    for data in payloads:
        assert rv.crc32(data) == zlib.crc32(data)
        for level in (0, 1, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
                co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
                comp = co.compress(data) + co.flush()
                assert rv.inflate_block(comp, len(data)) == data
                if len(data) > 100:
                    assert rv.inflate_block(comp, len(data) - 1) is None          # output too small
                    assert rv.inflate_block(comp[: len(comp) // 2], len(data)) != data  # truncated input
                n += 1
    assert n == len(payloads) * 16
    # every BGZF block of a BAM written by the generator
    d = cases.generate("c1_k1")
    raw = open(os.path.join(d, "S.bam"), "rb").read()
    p, blocks = 0, 0
    while p + 18 <= len(raw):
        bsize = struct.unpack_from("<H", raw, p + 16)[0] + 1
        crc, isize = struct.unpack_from("<II", raw, p + bsize - 8)
        want = zlib.decompress(raw[p + 18: p + bsize - 8], -15)
        got = rv.inflate_block(raw[p + 18: p + bsize - 8], isize)
        assert got == want and rv.crc32(got) == crc
        p += bsize
        blocks += 1
    assert blocks > 10
