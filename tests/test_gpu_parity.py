"""GPU parity tests proper: the CUDA path, driven through the C ABI (tests/tools/rv_dump.cpp --backend gpu
and the product CLI), against the oracle.  The oracle is the reference binary when oracle/_ref travelled to
the box, and in any case the committed golden vectors generated from it (tests/golden/make_golden.py).

Bar: integer / byte / index fields bit-exact; floating-point fields within 1e-9 relative.
"""
import gzip
import os

import numpy as np
import pytest

import cases
import dumpcmp
from conftest import ROOT, golden_path, run, unpack_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_cuda_path_matches_golden(built, name, tmp_path):
    c = cases.CASES[name]
    cases.generate(name)
    got = str(tmp_path / "gpu.txt")
    run(cases.dump_cmd(name, "gpu", got, c["stages"]))
    want = unpack_golden(name, "dump.txt", str(tmp_path / "golden.txt"))
    n, problems = dumpcmp.compare(want, got, c["exact_stages"])
    assert n > 0 or name == "edge_empty"
    assert not problems, problems[:5]


@pytest.mark.parametrize("env", [{"RV_G4_VARIANT": "3"}, {"RV_G4_VARIANT": "1", "RV_G4_RUN": "2"}, {"RV_NO_GATHER": "1"}])
def test_alternative_kernel_paths_match_golden(built, env, tmp_path):
    """The opt-in kernel variants must stay exact too: other launch shapes of the gather kernel and the path that sends
    every base through the literal walk and the sparse-observation list (RV_NO_GATHER=1, the debugging reference for
    the descriptor / gather path)."""
    name = "c5_k1"
    c = cases.CASES[name]
    cases.generate(name)
    got = str(tmp_path / "gpu.txt")
    run(cases.dump_cmd(name, "gpu", got, "C"), env=dict(os.environ, **env))
    want = unpack_golden(name, "dump.txt", str(tmp_path / "golden.txt"))
    n, problems = dumpcmp.compare(want, got, ["C."])
    assert n > 1000 and not problems, problems[:5]


@pytest.mark.parametrize("name", ["c1_k1", "c5_k1"])
def test_cuda_path_matches_reference_binary_run_here(built, ref_tools, name, tmp_path):
    if ref_tools is None:
        pytest.skip("oracle/_ref did not travel to this box")
    c = cases.CASES[name]
    d = cases.generate(name)
    want = str(tmp_path / "ref.txt")
    env = dict(os.environ, RV_DUMP=want, RV_DUMP_STAGES="C")
    run([ref_tools["ref_dump"]] + c["ref_args"](d), env=env)
    got = str(tmp_path / "gpu.txt")
    run(cases.dump_cmd(name, "gpu", got, "C"))
    n, problems = dumpcmp.compare(want, got, ["C."], rel_tol=0.0)
    assert n > 1000 and not problems, problems[:5]


def _tsv_lines(text):
    return sorted(l for l in text.splitlines() if l)


def _tsv_equal(a, b):
    ta, tb = a.split("\t"), b.split("\t")
    return len(ta) == len(tb) and all(dumpcmp.fields_equal(x, y, 2e-6) for x, y in zip(ta, tb))


@pytest.mark.parametrize("name", ["c1_k0", "c1_k1", "c5_k0", "c5_k1", "c3_k0", "c2_pileup_k0", "dedup_t", "dedup_t_F500",
                                  "c5_un_k1", "edge_nh_k1", "edge_nh_T80_k1"])
def test_cli_tsv_matches_reference_output(built, name, tmp_path):
    """End to end through the drop-in CLI: same flags as the reference, same TSV (sorted multiset of lines;
    %f-printed doubles may differ in the last printed digit when the oracle's -ffast-math quotient is off by an ulp)."""
    c = cases.CASES[name]
    d = cases.generate(name)
    out = str(tmp_path / "out.tsv")
    run([os.path.join(ROOT, "build", "rabbitvar_b200")] + c["ref_args"](d) + ["--out", out])
    got = _tsv_lines(open(out).read())
    with gzip.open(golden_path(name, "tsv"), "rt") as f:
        want = _tsv_lines(f.read())
    assert len(got) == len(want)
    bad = [(w, g) for w, g in zip(want, got) if w != g and not _tsv_equal(w, g)]
    assert not bad, bad[:3]


@pytest.mark.parametrize("name", sorted(cases.SOMATIC_CASES))
def test_cli_somatic_tsv_matches_reference_output(built, name, tmp_path):
    """Paired tumor | normal mode (BASELINE.json configs[1]) end to end through the drop-in CLI: the per-position
    join of the two samples, the classification labels and the 63-column lines of the reference binary."""
    c = cases.SOMATIC_CASES[name]
    d = cases.generate(name)
    out = str(tmp_path / "out.tsv")
    run([os.path.join(ROOT, "build", "rabbitvar_b200")] + c["ref_args"](d) + ["--out", out])
    got = _tsv_lines(open(out).read())
    with gzip.open(golden_path(name, "tsv"), "rt") as f:
        want = _tsv_lines(f.read())
    assert len(got) == len(want), (len(got), len(want))
    bad = [(w, g) for w, g in zip(want, got) if w != g and not _tsv_equal(w, g)]
    assert not bad, bad[:3]
    # the <out>.info side file: average coverage of the tumor and the normal sample
    info = [float(x) for x in open(out + ".info").read().split()]
    want_info = cases.SOMATIC_INFO[name]
    assert len(info) == 2 and all(abs(a - b) <= 2e-6 * max(1.0, abs(b)) for a, b in zip(info, want_info)), (info, want_info)


def test_tiled_bed_equals_per_tile_runs(built, tmp_path):
    """A tile is the parity unit: running N tiles in one batch must equal running them one at a time."""
    import rabbitvar_b200 as rv
    d = cases.generate("c1_k1")
    bam, fa = os.path.join(d, "S.bam"), os.path.join(d, "ref.fa")
    starts = [1301, 6301, 11301, 16301]
    ends = [6300, 11300, 16300, 21300]
    b = rv.HostBatch(bam, "chrS1", starts[0], ends[-1])
    ref = rv.fetch_ref(fa, "chrS1", 1, b.chr_len)
    lim = rv.default_limits(max_reads=b.n_reads + 16, max_read_bytes=b.pool_bytes + 64, max_ref_bases=len(ref) + 16)
    ctx = rv.Context(0, rv.default_params(), lim)
    ctx.set_reference(1, ref)
    ctx.push_reads(b)
    regs = b.make_regions(starts, ends)
    ctx.set_regions(regs)
    st_all = ctx.pileup()
    tabs_all = [ctx.fetch_tables(i) for i in range(4)]
    total = 0
    for i in range(4):
        one = b.make_regions(starts[i:i + 1], ends[i:i + 1])
        ctx.set_regions(one)
        st = ctx.pileup()
        total += st.n_aligned_bases
        counts, cov, first = ctx.fetch_tables(0)
        assert first == tabs_all[i][2]
        # field 7 keeps "the first (tp, q) recorded", which depends on atomic arrival order: compare its flags only
        assert np.array_equal(counts[..., :7], tabs_all[i][0][..., :7]) and np.array_equal(cov, tabs_all[i][1])
        assert np.array_equal(counts[..., 7] >> 24, tabs_all[i][0][..., 7] >> 24)
    assert total == st_all.n_aligned_bases
    ctx.close()


def test_pileup_is_deterministic_and_additive(built):
    """Size-independent properties of the atomics-based accumulation: repeating the launch gives identical
    tables (order independence), and the tables of two disjoint read halves add up to the table of the whole."""
    import ctypes as C
    import rabbitvar_b200 as rv
    d = cases.generate("c5_k1")
    bam, fa = os.path.join(d, "S.bam"), os.path.join(d, "ref.fa")
    b = rv.HostBatch(bam, "chrS5", 1301, 11300)
    ref = rv.fetch_ref(fa, "chrS5", 1, b.chr_len)
    lim = rv.default_limits(max_reads=b.n_reads + 16, max_read_bytes=b.pool_bytes + 64, max_ref_bases=len(ref) + 16)
    ctx = rv.Context(0, rv.default_params(move3=1, uniq_u=1), lim)
    ctx.set_reference(1, ref)
    ctx.push_reads(b)
    regs = b.make_regions([1301], [11300])
    ctx.set_regions(regs)
    ctx.pileup()
    c1, v1, _ = ctx.fetch_tables(0)
    ctx.pileup()
    c2, v2, _ = ctx.fetch_tables(0)
    assert np.array_equal(c1[..., :7], c2[..., :7]) and np.array_equal(v1, v2)
    assert np.array_equal(c1[..., 7] >> 24, c2[..., 7] >> 24)  # pstd/qstd flags (the recorded first value may differ)
    # halves
    n = b.n_reads
    lo, hi = regs[0].read_lo, regs[0].read_hi
    mid = (lo + hi) // 2
    parts = []
    for a, z in ((lo, mid), (mid, hi)):
        r = (rv.Region * 1)()
        C.memmove(r, regs, C.sizeof(rv.Region))
        r[0].read_lo, r[0].read_hi = a, z
        ctx.set_regions(r)
        ctx.pileup()
        parts.append(ctx.fetch_tables(0))
    assert np.array_equal(parts[0][0][..., :7] + parts[1][0][..., :7], c1[..., :7])
    assert np.array_equal(parts[0][1] + parts[1][1], v1)
    ctx.close()


def test_lazy_mode_gives_the_settled_results(built):
    """rv_set_lazy: rv_pileup / rv_score only enqueue; the first getter settles.  Statistics, variant records and
    the overflow error must be those of the settled calls."""
    import ctypes as C
    import rabbitvar_b200 as rv
    d = cases.generate("c1_k1")
    bam, fa = os.path.join(d, "S.bam"), os.path.join(d, "ref.fa")
    b = rv.HostBatch(bam, "chrS1", 1301, 21300)
    ref = rv.fetch_ref(fa, "chrS1", 1, b.chr_len)
    lim = rv.default_limits(max_reads=b.n_reads + 16, max_read_bytes=b.pool_bytes + 64, max_ref_bases=len(ref) + 16)
    ctx = rv.Context(0, rv.default_params(candidates_only=1), lim)
    ctx.set_reference(1, ref)
    ctx.push_reads(b)
    regs = b.make_regions([1301], [21300])
    ctx.set_regions(regs)
    st = ctx.pileup()
    ctx.score()
    vp, n = ctx.fetch_variants()
    want = bytes((C.c_char * (n * C.sizeof(rv.Variant))).from_address(C.addressof(vp.contents)))
    assert n > 0
    ctx.set_lazy(True)
    for _ in range(3):
        ctx.pileup_enqueue()
        ctx.score()
    assert ctx.n_variants() == n
    vp, n2 = ctx.fetch_variants()
    got = bytes((C.c_char * (n2 * C.sizeof(rv.Variant))).from_address(C.addressof(vp.contents)))
    st2 = rv.PileupStats()
    rv.lib().rv_get_pileup_stats(ctx._h, C.byref(st2))
    assert (st2.n_aligned_bases, st2.n_reads_kept, st2.n_events) == (st.n_aligned_bases, st.n_reads_kept, st.n_events)
    a, bms = ctx.kernel_ms()
    assert a > 0 and bms > 0
    ctx.set_lazy(False)
    ctx.close()
    assert n2 == n and sorted(got[i:i + C.sizeof(rv.Variant)] for i in range(0, len(got), C.sizeof(rv.Variant))) == \
        sorted(want[i:i + C.sizeof(rv.Variant)] for i in range(0, len(want), C.sizeof(rv.Variant)))
    # an overflow the enqueuing call could not report comes back from the settling call
    lim2 = rv.default_limits(max_reads=b.n_reads + 16, max_read_bytes=b.pool_bytes + 64, max_ref_bases=len(ref) + 16,
                             max_variants=4)
    ctx = rv.Context(0, rv.default_params(candidates_only=1), lim2)
    ctx.set_reference(1, ref)
    ctx.push_reads(b)
    ctx.set_regions(regs)
    ctx.set_lazy(True)
    ctx.pileup_enqueue()
    ctx.score()
    with pytest.raises(rv.RabbitVarError):
        ctx.sync()
    ctx.close()
    b.close()


def test_pipeline_equals_single_call(built):
    """rvh_pipeline_run (chunks of tiles on several worker contexts, ranged read uploads) must print exactly what one
    rvh_call_regions over all tiles prints."""
    import rabbitvar_b200 as rv
    d = cases.generate("c1_k1")
    bam, fa = os.path.join(d, "S.bam"), os.path.join(d, "ref.fa")
    starts = list(range(1301, 21301, 2500))
    ends = [s + 2499 for s in starts]
    b = rv.HostBatch(bam, "chrS1", starts[0], ends[-1])
    ref = rv.fetch_ref(fa, "chrS1", 1, b.chr_len)
    lim = rv.default_limits(max_reads=b.n_reads + 16, max_read_bytes=b.pool_bytes + 64, max_ref_bases=len(ref) + 16)
    params = rv.default_params()
    ctx = rv.Context(0, params, lim)
    regs = b.make_regions(starts, ends)
    want, tm = ctx.call_regions(b, regs, ref, 1, "S", "chrS1")
    ctx.close()
    assert tm.n_lines > 10
    for workers, chunk in ((1, 8), (3, 3), (4, 1)):
        pipe = rv.Pipeline(0, workers)
        got, tm2 = pipe.run(params, b, regs, chunk, ref, 1, "S", "chrS1")
        assert pipe.launch_count() > 0
        pipe.close()
        assert got == want
        assert tm2.n_aligned_bases == tm.n_aligned_bases and tm2.n_lines == tm.n_lines
    b.close()


def test_paired_pipeline_matches_reference_output(built):
    """rvh_pipeline_run_paired (chunks of tumor + normal tiles, two-range read uploads, several worker contexts) prints
    the reference binary's somatic-mode lines (the BED name column aside: the pipeline call carries no BED names)."""
    import ctypes as C
    import rabbitvar_b200 as rv
    name = "c2_somatic_bed_k0"
    d = cases.generate(name)
    tiles = [l.split() for l in open(os.path.join(d, "tiles.bed"))]
    starts, ends = [int(t[1]) for t in tiles], [int(t[2]) for t in tiles]
    bt = rv.HostBatch(os.path.join(d, "T.bam"), "chrS2", starts[0], ends[-1])
    bn = rv.HostBatch(os.path.join(d, "N.bam"), "chrS2", starts[0], ends[-1])
    n_t = bt.n_reads
    off = bt.append(bn)
    n_n = bn.n_reads
    bn.close()
    rt = bt.make_regions(starts, ends, 1200, 0, n_t)
    rn = bt.make_regions(starts, ends, 1200, off, n_n)
    n = len(tiles)
    regs = (rv.Region * (2 * n))()
    C.memmove(regs, rt, C.sizeof(rv.Region) * n)
    C.memmove(C.byref(regs, C.sizeof(rv.Region) * n), rn, C.sizeof(rv.Region) * n)
    ref = rv.fetch_ref(os.path.join(d, "ref.fa"), "chrS2", 1, bt.chr_len)
    params = rv.default_params(fisher=1, local_realign=0)
    with gzip.open(golden_path(name, "tsv"), "rt") as f:
        want = _tsv_lines(f.read())

    def strip_gene(lines):
        out = []
        for l in lines:
            t = l.split("\t")
            t[1] = ""
            out.append("\t".join(t))
        return sorted(out)

    want = strip_gene(want)
    for workers, chunk in ((1, 3), (3, 1)):
        pipe = rv.Pipeline(0, workers)
        tsv, tm = pipe.run(params, bt, regs, chunk, ref, 1, "T|N", "chrS2", paired=True)
        pipe.close()
        got = strip_gene(_tsv_lines(tsv))
        assert len(got) == len(want), (len(got), len(want))
        bad = [(w, g) for w, g in zip(want, got) if w != g and not _tsv_equal(w, g)]
        assert not bad, bad[:3]
    bt.close()


@pytest.mark.parametrize("key,scale", [("1", 0.2), ("1R", 0.1), ("2", 0.1), ("3", 0.1), ("4", 0.02), ("5", 0.1)])
def test_baseline_config_slice_matches_reference_binary(built, ref_tools, key, scale):
    """Every BASELINE.json config, same flags as tools/parity_configs.py runs them at full size (bench.py does, and
    records `parity_lines_differing`), here on a slice (>= 1 Mb for config 4): the drop-in CLI — parallel decode,
    GPU pipeline — against the reference binary built by oracle/Makefile, TSV as sorted multisets."""
    if ref_tools is None:
        pytest.skip("oracle/_ref did not travel to this box")
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import parity_configs as pc
    r = pc.run_config(key, scale, 1, 4)
    assert not r.get("error"), r
    assert r["ref_lines"] > 0 and r["cli_lines"] == r["ref_lines"], r
    assert r["parity_lines_differing"] == 0, r["examples"][:3]
    if key == "2":
        assert r["info_equal"]


def test_tile_blocks_concatenate_to_the_whole_run(built, tmp_path):
    """The multi-GPU split of bench.py (config 4: contiguous blocks of the tile list, one CLI process per block, text
    concatenated in block order) on one GPU: the blocks' TSV must add up to the TSV of the undivided run, and
    rvh_run_files (the CLI's loop behind the C ABI) must return the same text."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import parity_configs as pc
    import rabbitvar_b200 as rv
    from rabbitvar_b200 import shard
    d, length = pc.dataset("4", 0.02)
    args = pc.cli_args("4", d, length)
    bed = args[args.index("-i") + 1]
    tiles = [l.split() for l in open(bed)]
    whole = str(tmp_path / "whole.tsv")
    run([pc.CLI] + args + ["--th", "4", "--out", whole])
    w = shard.bai_tile_weights(os.path.join(d, "S.bam.bai"), 0, [(int(t[1]), int(t[2])) for t in tiles])
    parts = []
    for g, (lo, hi) in enumerate(shard.contiguous_blocks(w, 3)):
        sub = str(tmp_path / f"b{g}.bed")
        with open(sub, "w") as f:
            f.writelines("\t".join(t) + "\n" for t in tiles[lo:hi])
        a = list(args)
        a[a.index("-i") + 1] = sub
        out = str(tmp_path / f"b{g}.tsv")
        run([pc.CLI] + a + ["--th", "2", "--out", out])
        parts.append(open(out).read())
    assert "".join(parts) == open(whole).read()
    rc, text, _ = rv.run_files(rv.default_params(), os.path.join(d, "ref.fa"), os.path.join(d, "S.bam"),
                               [(t[0], int(t[1]), int(t[2]), t[3]) for t in tiles], sample="S", decode_threads=3)
    assert rc == 0 and text == open(whole).read()


def test_reference_binary_with_compiled_binding(built, ref_tools, tmp_path):
    """oracle/_ref/RabbitVar_b200 = the reference's own Launcher / RegionBuilder / CLI objects linked with
    oracle/ref_binding/simple_mode_b200.cpp in place of src/modes/simpleMode.cpp: its SimpleMode::process hands the
    regions to rvh_run_files (the C ABI).  Same flags, same files: its TSV must equal the stock reference binary's."""
    if ref_tools is None:
        pytest.skip("oracle/_ref did not travel to this box")
    bound = os.path.join(ROOT, "oracle", "_ref", "RabbitVar_b200")
    if not os.path.exists(bound):
        pytest.skip("oracle/_ref/RabbitVar_b200 was not built")
    d = cases.generate("c5_k1")
    bed = str(tmp_path / "two.bed")
    with open(bed, "w") as f:
        f.write("chrS5\t1301\t6300\tg1\nchrS5\t6301\t11300\tg2\n")
    args = ["-G", os.path.join(d, "ref.fa"), "-b", os.path.join(d, "S.bam"), "-N", "S", "-i", bed, "-c", "1", "-S", "2", "-E", "3",
            "-g", "4", "-f", "0.01", "-3", "-u", "--fisher", "--th", "2"]
    want_f, got_f = str(tmp_path / "ref.tsv"), str(tmp_path / "bound.tsv")
    run([ref_tools["RabbitVar"]] + args + ["--out", want_f])
    run([bound] + args + ["--out", got_f])
    want, got = _tsv_lines(open(want_f).read()), _tsv_lines(open(got_f).read())
    assert len(want) > 20 and len(got) == len(want)
    bad = [(w, g) for w, g in zip(want, got) if w != g and not _tsv_equal(w, g)]
    assert not bad, bad[:3]
