"""Seeded synthetic parity cases shared by the tests and tests/golden/make_golden.py.

Shapes follow BASELINE.json's five configs at sizes the reference finishes in seconds."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "build")
WORK = os.path.join(ROOT, "_work", "cases")


def _simple(chrom, region, extra=()):
    def f(d):
        return ["-G", os.path.join(d, "ref.fa"), "-b", os.path.join(d, "S.bam"), "-N", "S", "-f", "0.01",
                "-R", f"{chrom}:{region}"] + list(extra)
    return f


def _bed(chrom, bed, extra=()):
    def f(d):
        return ["-G", os.path.join(d, "ref.fa"), "-b", os.path.join(d, "S.bam"), "-N", "S", "-i",
                os.path.join(d, bed), "-c", "1", "-S", "2", "-E", "3", "-g", "4"] + list(extra)
    return f


# name -> synthgen args, reference CLI args, rv_dump args (our side), which stages are expected to match
CASES = {
    # cfg 1: single sample, simple mode, SNVs + a few indels; default -k 1 (CigarModifier + realignIndels): every stage exact
    "c1_k1": dict(gen=["--cfg", "1", "--len", "22600", "--depth", "100"], chrom="chrS1", region="1301-21300",
                  ref_args=_simple("chrS1", "1301-21300"), dump_args=[], stages="CRV", exact_stages=["C.", "R.", "V."]),
    # same data, -k 0: every stage (pileup, adjustMNP, scoring) must match
    "c1_k0": dict(gen=["--cfg", "1", "--len", "22600", "--depth", "100"], chrom="chrS1", region="1301-21300",
                  ref_args=_simple("chrS1", "1301-21300", ["-k", "0"]), dump_args=["--k", "0"], stages="CRV",
                  exact_stages=["C.", "R.", "V."]),
    # cfg 5: indel / soft-clip / MNV heavy with -3 -u
    "c5_k1": dict(gen=["--cfg", "5", "--len", "12600", "--depth", "100"], chrom="chrS5", region="1301-11300",
                  ref_args=_simple("chrS5", "1301-11300", ["-3", "-u"]), dump_args=["--three", "1", "--u", "1"],
                  stages="CRV", exact_stages=["C.", "R.", "V."]),
    "c5_k0": dict(gen=["--cfg", "5", "--len", "12600", "--depth", "100"], chrom="chrS5", region="1301-11300",
                  ref_args=_simple("chrS5", "1301-11300", ["-3", "-u", "-k", "0"]),
                  dump_args=["--three", "1", "--u", "1", "--k", "0"], stages="CRV", exact_stages=["C.", "R.", "V."]),
    # --UN with the CIGAR rewrite on: isReadsOverlap asks the record for its reference length AFTER CigarModifier
    # wrote the rewritten ops over the record's CIGAR (parseCigar.cpp:196-206, cigarModifier.cpp:386)
    "c5_un_k1": dict(gen=["--cfg", "5", "--len", "8600", "--depth", "100", "--seed", "21"], chrom="chrS5",
                     region="1301-7300", ref_args=_simple("chrS5", "1301-7300", ["-3", "--UN"]),
                     dump_args=["--three", "1", "--UN", "1"], stages="CRV", exact_stages=["C.", "R.", "V."]),
    # read bases 'N' (quality 2) in 30 % of the reads, a hard clip at one end of 20 %: the N skip of the M loop
    # (parseCigar.cpp:686-692), N inside soft clips (:1186), H ops (:655)
    "edge_nh_k1": dict(gen=["--cfg", "5", "--len", "10600", "--depth", "80", "--nbase-frac", "0.3", "--hardclip-frac", "0.2",
                            "--seed", "41"], chrom="chrS5", region="1301-9300",
                       ref_args=_simple("chrS5", "1301-9300", ["-3", "-u"]), dump_args=["--three", "1", "--u", "1"],
                       stages="CRV", exact_stages=["C.", "R.", "V."]),
    # -T 80 on the same data: reads are cut after 80 bases, so an insertion's anchor base can be a trimmed one and its
    # subtraction from the reference allele depends on an EARLIER read having created the entry (parseCigar.cpp:1474-1476)
    "edge_nh_T80_k1": dict(gen=["--cfg", "5", "--len", "10600", "--depth", "80", "--nbase-frac", "0.3", "--hardclip-frac", "0.2",
                                "--seed", "41"], chrom="chrS5", region="1301-9300",
                           ref_args=_simple("chrS5", "1301-9300", ["-3", "-u", "-T", "80"]),
                           dump_args=["--three", "1", "--u", "1", "--T", "80"],
                           stages="CRV", exact_stages=["C.", "R.", "V."]),
    # cfg 3: deep amplicon panel, low VAF, BED input (4-column BED => simple mode)
    "c3_k0": dict(gen=["--cfg", "3", "--len", "14600", "--depth", "2000", "--amplicons", "3"], chrom="chrS3",
                  bed="panel.bed", ref_args=_bed("chrS3", "panel.bed", ["-f", "0.005", "-k", "0"]),
                  dump_args=["--f", "0.005", "--k", "0"], stages="CV", exact_stages=["C.", "V."]),
    # cfg 2 shape: the tumor sample scored the way the somatic driver scores it (every covered position kept)
    "c2_pileup_k0": dict(gen=["--cfg", "1", "--len", "8600", "--depth", "150", "--seed", "99"], chrom="chrS1",
                         region="1301-7300", ref_args=_simple("chrS1", "1301-7300", ["-k", "0", "-p", "--fisher"]),
                         dump_args=["--k", "0", "--p", "1", "--fisher", "1"], stages="CV", exact_stages=["C.", "V."]),
    # -t (recordPreprocessor.cpp:153-176): half of the fragments are single-end records, whose POS-RNEXT-PNEXT key
    # ("POS-*-0") makes every later single-end record of the same start a duplicate
    "dedup_t": dict(gen=["--cfg", "1", "--len", "12600", "--depth", "150", "--single-frac", "0.5", "--seed", "11"],
                    chrom="chrS1", region="1301-11300", ref_args=_simple("chrS1", "1301-11300", ["-t"]),
                    dump_args=["--t", "1"], stages="CRV", exact_stages=["C.", "R.", "V."]),
    # -t with -F 0x500: paired records flagged 0x4 survive the flag filter and take the POS-CIGAR key
    "dedup_t_F500": dict(gen=["--cfg", "5", "--len", "8600", "--depth", "400", "--single-frac", "0.2",
                              "--unmapped-frac", "0.6", "--seed", "12"],
                         chrom="chrS5", region="1301-7300",
                         ref_args=_simple("chrS5", "1301-7300", ["-t", "-F", "0x500", "-k", "0"]),
                         dump_args=["--t", "1", "--F", "500", "--k", "0"], stages="CRV",
                         exact_stages=["C.", "R.", "V."]),
    # empty / ragged: a region with no reads at all and a region at the contig edge of the read cloud
    "edge_empty": dict(gen=["--cfg", "1", "--len", "8600", "--depth", "20", "--seed", "5"], chrom="chrS1",
                       region="10-1250", ref_args=_simple("chrS1", "10-1250"), dump_args=[], stages="CRV",
                       exact_stages=["C.", "R.", "V."]),
}


def _somatic(chrom, region=None, bed=None, extra=()):
    def f(d):
        a = ["-G", os.path.join(d, "ref.fa"), "-b", os.path.join(d, "T.bam") + "|" + os.path.join(d, "N.bam"), "-N", "T|N",
             "-f", "0.01", "--fisher"]
        if region:
            a += ["-R", f"{chrom}:{region}"]
        else:
            a += ["-i", os.path.join(d, bed), "-c", "1", "-S", "2", "-E", "3", "-g", "4"]
        return a + list(extra)
    return f


# paired tumor | normal cases (BASELINE.json configs[1]): only the CLI's TSV is compared (the stage dumps of
# oracle/ref_dump cover one sample at a time and are exercised by the cases above)
SOMATIC_CASES = {
    "c2_somatic_k1": dict(gen=["--cfg", "2", "--len", "22600", "--depth", "120", "--depth-n", "60"], chrom="chrS2",
                          ref_args=_somatic("chrS2", region="1301-21300")),
    "c2_somatic_bed_k0": dict(gen=["--cfg", "2", "--len", "32600", "--depth", "80", "--depth-n", "50", "--seed", "31"],
                              chrom="chrS2", ref_args=_somatic("chrS2", bed="tiles.bed", extra=["-k", "0"])),
}


# contents of the reference's <out>.info side file for the cases above (tests/golden/make_golden.py prints them)
SOMATIC_INFO = {"c2_somatic_k1": (114.754443, 56.724382), "c2_somatic_bed_k0": (78.601912, 49.945043)}


def dataset_dir(name):
    return os.path.join(WORK, name)


def generate(name):
    d = dataset_dir(name)
    marker = os.path.join(d, "meta.txt")
    if not os.path.exists(marker):
        os.makedirs(d, exist_ok=True)
        gen = (CASES.get(name) or SOMATIC_CASES[name])["gen"]
        subprocess.run([os.path.join(BUILD, "synthgen")] + gen + ["--out", d], check=True, stderr=subprocess.DEVNULL)
    return d


def dump_cmd(name, backend, out, stages):
    c = CASES[name]
    d = dataset_dir(name)
    cmd = [os.path.join(BUILD, "rv_dump"), "--backend", backend, "--fasta", os.path.join(d, "ref.fa"), "--bam",
           os.path.join(d, "S.bam"), "--chr", c["chrom"], "--out", out, "--stages", stages]
    if "region" in c:
        cmd += ["--region", c["region"]]
    else:
        cmd += ["--bed", os.path.join(d, c["bed"])]
    return cmd + c["dump_args"]
