"""Host decode path: BGZF/BAM/BAI/FAI reader against the generator's own bookkeeping."""
import os

import numpy as np
import pytest

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


def test_block_decoder_second_level_tables_and_damaged_streams(built):
    """Codewords longer than the first-level tables (11 bits literal/length, 8 bits distance) go through second-level
    tables: skewed byte distributions under Z_HUFFMAN_ONLY give 12-15 bit literal codes, rare far copies between long
    runs give long distance codes.  Then streams with flipped bits: whatever zlib accepts must come out identical, and
    nothing may be written beyond the output size asked for (the binding checks the returned length)."""
    import random
    import zlib
    import rabbitvar_b200 as rv
    rnd = random.Random(11)
    payloads = []
    for lam in (0.02, 0.05, 0.1, 0.3):  # geometric-ish byte values: a few frequent symbols, a long tail of rare ones
        payloads.append(bytes(min(255, int(rnd.expovariate(lam))) for _ in range(120000)))
    for _ in range(4):  # long runs with rare copies from far back at many different distances
        buf = bytearray(bytes(rnd.getrandbits(8) for _ in range(3000)))
        while len(buf) < 150000:
            if rnd.random() < 0.9:
                buf += bytes([rnd.choice((0, 0, 0, 7))]) * rnd.randint(3, 300)
            else:
                d = rnd.randint(1, min(len(buf), 32768))
                ln = rnd.randint(3, 40)
                buf += buf[len(buf) - d: len(buf) - d + ln]
        payloads.append(bytes(buf))
    n_long = 0
    streams = []
    for data in payloads:
        for level, strategy in ((1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY),
                                (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE), (6, zlib.Z_FILTERED)):
            co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
            comp = co.compress(data) + co.flush()
            assert rv.inflate_block(comp, len(data)) == data
            streams.append((comp, data))
            n_long += 1
    assert n_long == len(payloads) * 6
    # damaged copies
    agree = 0
    for k in range(400):
        comp, data = streams[k % len(streams)]
        bad = bytearray(comp)
        for _ in range(rnd.randint(1, 3)):
            bad[rnd.randrange(len(bad))] ^= 1 << rnd.randrange(8)
        bad = bytes(bad)
        try:
            d = zlib.decompressobj(-15)
            want = d.decompress(bad) + d.flush()
            z_ok = d.eof and len(want) == len(data)
        except zlib.error:
            z_ok = False
        got = rv.inflate_block(bad, len(data))
        if z_ok:
            assert got == want
            agree += 1
        else:
            assert got is None or len(got) <= len(data)
    assert agree >= 0


def _python_bam_records(path):
    """An independent BAM reader for the test below: Python's gzip module (multi-member = BGZF) and the record layout
    of the SAM/BAM specification section 4.2, nothing from the repo's C++ reader."""
    import gzip
    import struct
    raw = gzip.open(path, "rb").read()
    assert raw[:4] == b"BAM\1"
    l_text, = struct.unpack_from("<i", raw, 4)
    o = 8 + l_text
    n_ref, = struct.unpack_from("<i", raw, o)
    o += 4
    refs = []
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<i", raw, o)
        name = raw[o + 4:o + 4 + l_name - 1].decode()
        l_ref, = struct.unpack_from("<i", raw, o + 4 + l_name)
        refs.append((name, l_ref))
        o += 8 + l_name
    recs = []
    while o < len(raw):
        bs, = struct.unpack_from("<i", raw, o)
        tid, pos, l_qname, mapq, _bin, n_cigar, flag, l_seq, mtid, mpos, _tlen = struct.unpack_from("<iiBBHHHiiii", raw, o + 4)
        p = o + 36 + l_qname
        cigar = struct.unpack_from("<%dI" % n_cigar, raw, p)
        tail = raw[p:p + 4 * n_cigar + (l_seq + 1) // 2 + l_seq]  # cigar | packed bases | qualities
        ref_len = sum(c >> 4 for c in cigar if (c & 15) in (0, 2, 3, 7, 8))
        a = p + len(tail)
        end = o + 4 + bs
        nm = -1
        while a < end:  # the aux fields, for NM
            tag, ty = raw[a:a + 2], raw[a + 2:a + 3]
            a += 3
            size = {b"A": 1, b"c": 1, b"C": 1, b"s": 2, b"S": 2, b"i": 4, b"I": 4, b"f": 4}.get(ty)
            if size is not None:
                if tag == b"NM":
                    nm = int.from_bytes(raw[a:a + size], "little", signed=ty in (b"c", b"s", b"i"))
                a += size
            elif ty in (b"Z", b"H"):
                a = raw.index(b"\0", a) + 1
            elif ty == b"B":
                sub = raw[a:a + 1]
                cnt, = struct.unpack_from("<i", raw, a + 1)
                a += 5 + cnt * {b"c": 1, b"C": 1, b"s": 2, b"S": 2, b"i": 4, b"I": 4, b"f": 4}[sub]
            else:
                raise AssertionError("aux type %r" % ty)
        recs.append(dict(tid=tid, pos=pos, mpos=mpos, flag=flag, l_seq=l_seq, n_cigar=n_cigar, mapq=mapq, mtid=mtid,
                         end=pos + (ref_len if ref_len else 1), nm=nm, tail=tail))
        o = end
    return refs, recs


@pytest.mark.parametrize("name,bam,chrom", [("c5_k1", "S.bam", "chrS5"), ("edge_nh_k1", "S.bam", None)])
def test_loader_matches_an_independent_python_reader(built, name, bam, chrom):
    """rvh_load_bam (BAI query, the repo's inflate, record decode, pool layout) against a reader written from the BAM
    specification with Python's gzip + struct: every field of every record, the cigar | bases | qualities bytes, and
    htslib's overlap rule on a sub-range served through the index."""
    d = cases.generate(name)
    refs, recs = _python_bam_records(os.path.join(d, bam))
    chrom = chrom or refs[0][0]
    tid = [r[0] for r in refs].index(chrom)
    dt = np.dtype([("pos", "<i4"), ("mpos", "<i4"), ("off", "<u4"), ("l_seq", "<i4"), ("flag", "<u2"),
                   ("n_cigar", "<u2"), ("nm", "<i2"), ("mapq", "u1"), ("same", "u1"), ("end", "<i4"), ("mtid", "<i4")])
    length = refs[tid][1]
    mine = [r for r in recs if r["tid"] == tid]
    mid, last = mine[len(mine) // 2]["pos"] + 1, mine[-1]["pos"] + 1
    for lo, hi in ((1, length), (mid, mid + 700), (last + 3, min(length, last + 40)), (mine[0]["pos"] + 1, mine[0]["pos"] + 1)):
        b = rv.HostBatch(os.path.join(d, bam), chrom, lo, hi)
        assert b.chr_len == length
        got = b.reads_numpy().view(dt)
        pool = b.pool_numpy()
        want = [r for r in recs if r["tid"] == tid and r["pos"] < hi and r["end"] > lo - 1]  # 0-based pos < end, endpos > beg
        assert len(got) == len(want) > 0
        for g, w in zip(got, want):
            assert (g["pos"], g["mpos"], g["flag"], g["l_seq"], g["n_cigar"], g["mapq"], g["end"], g["nm"]) == \
                (w["pos"] + 1, w["mpos"] + 1, w["flag"], w["l_seq"], w["n_cigar"], w["mapq"], w["end"], w["nm"])
            assert g["same"] == (1 if w["tid"] == w["mtid"] else 0) and g["mtid"] == w["mtid"]
            o = int(g["off"]) * 16
            assert pool[o:o + len(w["tail"])].tobytes() == w["tail"]
        b.close()


def _reg2bin(beg, end):
    """SAM specification section 5.3 (C code of the specification, restated): bin of the 0-based half-open [beg, end)."""
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


@pytest.mark.parametrize("name,bam", [("c5_k1", "S.bam"), ("c1_k1", "S.bam")])
def test_bai_written_by_the_generator_follows_the_specification(built, name, bam):
    """The .bai the loader's queries rest on is written by the repo's own code (tools/synthgen via csrc/io/bamio.hpp), so
    a reader and a writer that agree with each other but not with the format would go unnoticed.  Here the index is
    parsed from the SAM specification's section 5.2 layout with struct, the records and their virtual offsets come from a
    BGZF walk with zlib, and the two must fit: every record lies inside a chunk of the bin reg2bin() gives it (the bin
    field of the record says the same), and the linear index holds, per 16 kb window, the virtual offset of the first
    record that overlaps it."""
    import struct
    import zlib
    d = cases.generate(name)
    raw = open(os.path.join(d, bam), "rb").read()
    # BGZF walk: uncompressed stream + where each block starts in both coordinates
    blocks, data, p = [], bytearray(), 0
    while p + 18 <= len(raw):
        bsize = struct.unpack_from("<H", raw, p + 16)[0] + 1
        blocks.append((p, len(data)))
        data += zlib.decompress(raw[p + 18: p + bsize - 8], -15)
        p += bsize
    data = bytes(data)
    ustarts = [u for _, u in blocks]

    def voffset(u):  # virtual offset of uncompressed position u
        import bisect
        i = bisect.bisect_right(ustarts, u) - 1
        while i + 1 < len(blocks) and blocks[i + 1][1] == u:  # an empty block: the position belongs to the next one
            i += 1
        return (blocks[i][0] << 16) | (u - blocks[i][1])

    assert data[:4] == b"BAM\1"
    l_text, = struct.unpack_from("<i", data, 4)
    o = 8 + l_text
    n_ref, = struct.unpack_from("<i", data, o)
    o += 4
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<i", data, o)
        o += 8 + l_name
    recs = []
    while o < len(data):
        bs, = struct.unpack_from("<i", data, o)
        tid, pos, l_qname, mapq, bin_, n_cigar, flag, l_seq = struct.unpack_from("<iiBBHHHi", data, o + 4)
        cigar = struct.unpack_from("<%dI" % n_cigar, data, o + 36 + l_qname)
        ref_len = sum(c >> 4 for c in cigar if (c & 15) in (0, 2, 3, 7, 8))
        end = pos + (ref_len if ref_len and not flag & 4 else 1)
        recs.append((tid, pos, end, bin_, voffset(o), o, o + 4 + bs))
        o += 4 + bs
    assert len(recs) > 100
    # the index, section 5.2
    bai = open(os.path.join(d, bam + ".bai"), "rb").read()
    assert bai[:4] == b"BAI\1"
    n, = struct.unpack_from("<i", bai, 4)
    assert n == n_ref
    q = 8
    index = []
    for _ in range(n):
        n_bin, = struct.unpack_from("<i", bai, q)
        q += 4
        bins = {}
        for _ in range(n_bin):
            b, n_chunk = struct.unpack_from("<Ii", bai, q)
            q += 8
            bins[b] = [struct.unpack_from("<QQ", bai, q + 16 * k) for k in range(n_chunk)]
            q += 16 * n_chunk
        n_intv, = struct.unpack_from("<i", bai, q)
        q += 4
        lin = list(struct.unpack_from("<%dQ" % n_intv, bai, q))
        q += 8 * n_intv
        index.append((bins, lin))
    assert q == len(bai) or q + 8 == len(bai)  # (optional n_no_coor)
    first_in_window = {}
    cstart = {c: u for c, u in blocks}

    def upos(v):  # uncompressed position of a virtual offset (the end of a block and the start of the next are one place)
        return cstart[v >> 16] + (v & 0xffff)

    for tid, pos, end, bin_, v0, u0, u1 in recs:
        if tid < 0:
            continue
        want_bin = _reg2bin(pos, end)
        assert bin_ == want_bin
        bins, lin = index[tid]
        assert want_bin in bins
        assert any(upos(c0) <= u0 and u1 <= upos(c1) for c0, c1 in bins[want_bin]), (tid, pos, hex(v0), bins[want_bin][:3])
        for w in range(pos >> 14, ((end - 1) >> 14) + 1):
            first_in_window.setdefault((tid, w), v0)
    for (tid, w), v in first_in_window.items():
        lin = index[tid][1]
        assert w < len(lin) and lin[w] == v, (tid, w, hex(lin[w]) if w < len(lin) else None, hex(v))


def test_cli_narrows_the_visible_devices(built):
    """`--device d --gpus n` keeps the entries [d, d + n) of CUDA_VISIBLE_DEVICES (or of 0, 1, 2, ... when it is unset)
    before CUDA starts: a fresh process initialises every device it can see.  Checked on the decode-only path, which
    needs no GPU."""
    import subprocess
    d = cases.generate("c1_k1")
    cli = os.path.join(cases.ROOT, "build", "rabbitvar_b200")
    base = [cli, "-G", os.path.join(d, "ref.fa"), "-b", os.path.join(d, "S.bam"), "-R", "chrS1:1301-2300", "--decode-only",
            "--out", os.path.join(d, "narrow.tsv")]
    for vis, dev, gpus, want in (("0,1,2,3", 1, 2, "1,2"), ("0,1,2,3", 3, 1, "3"), ("GPU-aa, GPU-bb", 1, 1, "GPU-bb"),
                                 (None, 2, 1, "2"), (None, 0, 2, "0,1"), ("5", 0, 1, "5"), ("0,1", 1, 4, "1")):
        env = {k: v for k, v in os.environ.items() if k != "CUDA_VISIBLE_DEVICES"}
        if vis is not None:
            env["CUDA_VISIBLE_DEVICES"] = vis
        r = subprocess.run(base + ["--device", str(dev), "--gpus", str(gpus)], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr[-300:]
        assert f"CUDA_VISIBLE_DEVICES of this run: {want}\n" in r.stdout, (vis, dev, gpus, r.stdout[-400:])
