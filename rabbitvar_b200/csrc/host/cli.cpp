// cli.cpp — command-line front end that keeps the reference's CLI, region/BED handling and TSV output
// for the pileup-and-score path (reference src/Launcher.cpp:295-496 cmdParse, src/RegionBuilder.cpp:45-109
// buildRegions, :179-200 buildRegionFromConfiguration, src/modes/simpleMode.cpp:210-387 SimpleMode::process).
//
// The OpenMP region loop of the reference becomes region batches sharded over GPUs: the ordered region
// list is cut into contiguous blocks, block g goes to GPU g (one host thread + one rv_ctx per GPU), and the
// host concatenates the TSV blocks in region order (replaces the `omp critical` write, simpleMode.cpp:339).
// No collective sits on the path.
//
// Supported: simple mode (-b one BAM) and paired somatic mode (-b 'tumor.bam|normal.bam', somaticMode.cpp), -R or
// -i BED with -c/-S/-E/-g, the scoring/filter flags below.
#include "../../../include/rabbitvar_b200.h"
#include "pipeline.hpp"
#include <thread>
#include <fstream>
#include <sstream>
#include <iostream>
#include <map>

using namespace rvhost;

struct Cli {
  std::string fasta, bam, bam2, region, bed, out = "./out.txt", sample, delim = "\t";
  int c_col = 2, S_col = 6, E_col = 7, g_col = 12;  // DEFAULT_BED_ROW_FORMAT (Launcher.cpp:21), 0-based after -1
  bool c_set = false, S_set = false, E_set = false, g_set = false, zero_based = false;
  int nucl_ext = 0, ref_ext = 1200, gpus = 1, threads = 1, batch_regions = 256;
  rv_params P;
};

static void usage() {
  fprintf(stderr,
          "usage: rabbitvar_b200 -G ref.fa -b in.bam (-R chr:start-end | -i regions.bed -c 1 -S 2 -E 3 -g 4) [options]\n"
          "  -N name  -f freq  -k 0|1  -3  -u  --UN  -p  -t  --fisher  -q phred  -m mismatches  -X ext  -P pos  -r minr\n"
          "  -B minbias  -Q mapq  -o qratio  -O mapq  -V lofreq  -M minmatch  -T trim  -F hexfilter  -x extend  -Y refext\n"
          "  -z  --th n  --gpus n  --out file\n");
}

static bool parse(int argc, char** argv, Cli& c) {
  rv_default_params(&c.P);
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    std::string v;
    size_t eq = a.find('=');
    bool has_inline = a.rfind("--", 0) == 0 && eq != std::string::npos;
    if (has_inline) { v = a.substr(eq + 1); a = a.substr(0, eq); }
    auto val = [&]() -> std::string {
      if (has_inline) return v;
      if (i + 1 >= argc) { fprintf(stderr, "option needs value: %s\n", a.c_str()); exit(1); }
      return argv[++i];
    };
    if (a == "-G" || a == "--Genome_fasta") c.fasta = val();
    else if (a == "-b" || a == "--in_bam") c.bam = val();
    else if (a == "-R" || a == "--Region") c.region = val();
    else if (a == "-i" || a == "--bed") c.bed = val();
    else if (a == "--out") c.out = val();
    else if (a == "-N" || a == "--Name") c.sample = val();
    else if (a == "-d" || a == "--delemiter") c.delim = val();
    else if (a == "-c" || a == "--column") { c.c_col = atoi(val().c_str()) - 1; c.c_set = true; }
    else if (a == "-S" || a == "--region_start") { c.S_col = atoi(val().c_str()) - 1; c.S_set = true; }
    else if (a == "-E" || a == "--region_end") { c.E_col = atoi(val().c_str()) - 1; c.E_set = true; }
    else if (a == "-g" || a == "--gene_name") { c.g_col = atoi(val().c_str()) - 1; c.g_set = true; }
    else if (a == "-s" || a == "--seg_start" || a == "-e" || a == "--seg_end") val();
    else if (a == "-f" || a == "--allele_fre") c.P.freq = atof(val().c_str());
    else if (a == "-k" || a == "--local_realig") c.P.local_realign = atoi(val().c_str()) == 1;
    else if (a == "-3" || a == "--3-prime") c.P.move3 = 1;
    else if (a == "-u" || a == "--uni") c.P.uniq_u = 1;
    else if (a == "--UN") c.P.uniq_un = 1;
    else if (a == "-t" || a == "--dedup") c.P.dedup = 1;
    else if (a == "-p" || a == "--pileup") { c.P.pileup = 1; }
    else if (a == "--fisher") c.P.fisher = 1;
    else if (a == "-z" || a == "--zero_based") c.zero_based = true;
    else if (a == "-q" || a == "--phred_score") c.P.goodq = atof(val().c_str());
    else if (a == "-m" || a == "--mismatch") c.P.mismatch = atoi(val().c_str());
    else if (a == "-X" || a == "--extension") c.P.vext = atoi(val().c_str());
    else if (a == "-P" || a == "--Position") c.P.read_pos_filter = atoi(val().c_str());
    else if (a == "-r" || a == "--minimum_reads") c.P.minr = atoi(val().c_str());
    else if (a == "-B" || a == "--min") c.P.min_bias_reads = atoi(val().c_str());
    else if (a == "-Q" || a == "--Quality") c.P.mapping_quality = atoi(val().c_str());
    else if (a == "-o" || a == "--Qratio") c.P.qratio = atof(val().c_str());
    else if (a == "-O" || a == "--MapQ") c.P.mapq = atof(val().c_str());
    else if (a == "-V" || a == "--freq") c.P.lofreq = atof(val().c_str());
    else if (a == "-M" || a == "--Min_macth") c.P.minmatch = atoi(val().c_str());
    else if (a == "-T" || a == "--trim") c.P.trim_bases_after = atoi(val().c_str());
    else if (a == "-I" || a == "--Indel_size") c.P.indelsize = atoi(val().c_str());
    else if (a == "-F" || a == "--Filter") c.P.samfilter = (int)strtol(val().c_str(), NULL, 16);
    else if (a == "-x" || a == "--numcl_extend") c.nucl_ext = atoi(val().c_str());
    else if (a == "-Y" || a == "--ref-extension") c.ref_ext = atoi(val().c_str());
    else if (a == "--th") c.threads = std::max(1, atoi(val().c_str()));
    else if (a == "--gpus") c.gpus = std::max(1, atoi(val().c_str()));
    else if (a == "--batch-regions") c.batch_regions = std::max(1, atoi(val().c_str()));
    else if (a == "--auto_resize" || a == "-y" || a == "--verbose" || a == "--chimeric" || a == "--deldupvar") {}
    else if (a == "-Z" || a == "--downsample") {
      // recordPreprocessor.cpp:133 drops records by rand(): "random and non-reproducible" (Launcher.cpp:355)
      fprintf(stderr, "-Z (random downsampling) is not supported: its output is not reproducible by definition\n");
      exit(1);
    }
    else if (a == "-H" || a == "--help") { usage(); exit(0); }
    else if (a == "--version") { printf("rabbitvar_b200 0.1 (ABI %d)\n", rv_abi_version()); exit(0); }
    else { fprintf(stderr, "unrecognized option: %s\n", a.c_str()); usage(); exit(1); }
  }
  if (c.P.pileup) { c.P.freq = -1; c.P.minr = 0; }  // Launcher.cpp:455-459
  c.P.candidates_only = c.P.pileup ? 0 : 1;         // simple-mode output prints only passing variants
  if (c.fasta.empty() || c.bam.empty()) { usage(); return false; }
  size_t bar = c.bam.find('|');
  if (bar != std::string::npos) {  // BamNames, Configuration.h: "bam1|bam2" selects the somatic mode
    c.bam2 = c.bam.substr(bar + 1);
    c.bam = c.bam.substr(0, bar);
    if (c.bam.empty() || c.bam2.empty()) { usage(); return false; }
    // getSampleNamesSomatic (Launcher.cpp:250-279) is only used together with -R (Launcher.cpp:46-50): it prints
    // the first of the "T|N" names; with a BED file the -N string is printed as given
    size_t sb = c.sample.find('|');
    if (sb != std::string::npos && !c.region.empty()) c.sample = c.sample.substr(0, sb);
  }
  return true;
}

static std::vector<std::string> split(const std::string& s, const std::string& delims) {
  std::vector<std::string> out;
  size_t prev = 0, next;
  while ((next = s.find_first_of(delims, prev)) != std::string::npos) {
    if (next - prev != 0) out.push_back(s.substr(prev, next - prev));
    prev = next + 1;
  }
  if (prev < s.size()) out.push_back(s.substr(prev));
  return out;
}

// RegionBuilder (src/RegionBuilder.cpp:45-109, :179-200)
static bool build_regions(const Cli& c, const rvio::BamHeader& hdr, std::vector<RegionSpec>* out) {
  auto correct_chr = [&](std::string chr) {
    if (hdr.tid_of(chr) < 0) {
      if (chr.rfind("chr", 0) == 0) chr = chr.substr(3);
      else chr = "chr" + chr;
    }
    return chr;
  };
  if (!c.region.empty()) {
    std::vector<std::string> sp = split(c.region, ":");
    if (sp.size() < 2) return false;
    RegionSpec r;
    r.chr = correct_chr(sp[0]);
    r.gene = sp.size() < 3 ? r.chr : sp[2];
    std::vector<std::string> range = split(sp[1], "-");
    auto num = [](std::string s) { s.erase(std::remove(s.begin(), s.end(), ','), s.end()); return atoi(s.c_str()); };
    r.start = num(range[0]);
    r.end = range.size() < 2 ? r.start : num(range[1]);
    r.start -= c.nucl_ext;
    r.end += c.nucl_ext;
    if (c.zero_based && r.start < r.end) r.start++;
    if (r.start > r.end) r.start = r.end;
    out->push_back(r);
    return true;
  }
  std::ifstream in(c.bed);
  if (!in) return false;
  std::string line;
  while (std::getline(in, line)) {
    if (line.rfind("#", 0) == 0 || line.rfind("browser", 0) == 0 || line.rfind("track", 0) == 0) continue;
    std::vector<std::string> col = split(line, c.delim);
    int cc = c.c_col, sc = c.S_col, ec = c.E_col, gc = c.g_col;
    if ((int)col.size() <= std::max(cc, std::max(sc, ec))) continue;
    RegionSpec r;
    r.chr = correct_chr(col[cc]);
    int cds_start = atoi(col[sc].c_str()), cds_end = atoi(col[ec].c_str());
    r.gene = gc < (int)col.size() ? col[gc] : r.chr;
    // thick start/end default to the region columns when only -S/-E are given (Launcher.cpp:417-424)
    int ts = cds_start, te = cds_end;
    if (cds_start > te) continue;
    ts -= c.nucl_ext;
    te += c.nucl_ext;
    if (c.zero_based && ts < te) ts++;
    r.start = ts;
    r.end = te;
    out->push_back(r);
  }
  return true;
}

struct Block {
  std::vector<RegionSpec> specs;  // same contig, ascending
  std::string tsv;
  int64_t bases = 0, reads = 0, lines = 0;
  int64_t cov_sum[2] = {0, 0}, cov_pos[2] = {0, 0};
  double pileup_ms = 0, score_ms = 0;
  std::string err;
};

static void worker(const Cli& c, int device, std::vector<Block*> blocks) {
  rvio::BamReader bam;
  rvio::BaiIndex bai;
  rvio::Fasta fa;
  if (!bam.open(c.bam) || !bai.load(c.bam + ".bai") || !fa.open(c.fasta)) {
    for (Block* b : blocks) b->err = "cannot open BAM/BAI/FASTA";
    return;
  }
  const bool somatic = !c.bam2.empty();
  rvio::BamReader bamN;
  rvio::BaiIndex baiN;
  if (somatic && (!bamN.open(c.bam2) || !baiN.load(c.bam2 + ".bai"))) {
    for (Block* b : blocks) b->err = "cannot open the second BAM/BAI";
    return;
  }
  rv_ctx* ctx = NULL;
  rv_limits L;
  rv_default_limits(&L);
  int64_t cap_reads = 0, cap_bytes = 0, cap_pos = 0, cap_ref = 0;
  for (Block* blk : blocks) {
    try {
      const std::string& chr = blk->specs[0].chr;
      int tid = bam.header().tid_of(chr);
      if (tid < 0) { blk->err = "contig not in BAM: " + chr; continue; }
      int32_t chr_len = bam.header().lens[tid];
      int32_t smin = blk->specs[0].start, smax = blk->specs[0].end;
      for (auto& s : blk->specs) { smin = std::min(smin, s.start); smax = std::max(smax, s.end); }
      ReadBatch batch;
      load_span(bam, bai, tid, smin, smax, &batch);
      std::vector<rv_region> regs;
      make_regions(batch, blk->specs, chr_len, c.ref_ext, c.nucl_ext, &regs);
      if (somatic) {  // the same tiles of the normal sample follow the tumor's (one_region_run_somt)
        int tidN = bamN.header().tid_of(chr);
        if (tidN < 0) { blk->err = "contig not in the second BAM: " + chr; continue; }
        ReadBatch batchN;
        load_span(bamN, baiN, tidN, smin, smax, &batchN);
        std::vector<rv_region> regsN;
        make_regions(batchN, blk->specs, chr_len, c.ref_ext, c.nucl_ext, &regsN);
        const int64_t off = append_batch(batch, batchN);
        for (auto& r : regsN) { r.read_lo += off; r.read_hi += off; regs.push_back(r); }
      }
      int32_t ref_lo = std::max(1, smin - c.ref_ext - c.nucl_ext - 100);
      int32_t ref_hi = std::min(chr_len, smax + c.ref_ext + c.nucl_ext + 100);
      std::string refseq;
      fa.fetch(chr, ref_lo, ref_hi, &refseq);
      for (auto& ch : refseq) ch = (char)toupper((unsigned char)ch);
      int64_t npos = 0;
      for (auto& r : regs) npos += r.end - r.start + 1 + 2 * L.halo;
      if (!ctx || (int64_t)batch.reads.size() > cap_reads || (int64_t)batch.pool.size() > cap_bytes || npos > cap_pos ||
          (int64_t)refseq.size() > cap_ref) {
        if (ctx) rv_destroy(ctx);
        cap_reads = std::max<int64_t>(cap_reads, (int64_t)batch.reads.size() * 5 / 4 + 1024);
        cap_bytes = std::max<int64_t>(cap_bytes, (int64_t)batch.pool.size() * 5 / 4 + 4096);
        cap_pos = std::max<int64_t>(cap_pos, npos * 5 / 4 + 1024);
        cap_ref = std::max<int64_t>(cap_ref, (int64_t)refseq.size() * 5 / 4 + 1024);
        L.max_reads = cap_reads;
        L.max_read_bytes = cap_bytes;
        L.max_positions = cap_pos;
        L.max_regions = (int32_t)std::max<size_t>(regs.size(), (size_t)c.batch_regions * (somatic ? 2 : 1)) + 1;
        L.max_events = std::max<int64_t>(1 << 16, cap_reads * 4);
        L.max_variants = (somatic ? 3 : 1) * cap_pos + 1024;
        L.max_patch = std::max<int64_t>(1 << 16, cap_reads);
        L.max_ref_bases = cap_ref;
        int rc = rv_create(&ctx, device, &c.P, &L);
        if (rc != RV_OK) {
          blk->err = std::string("rv_create: ") + (ctx ? rv_last_error(ctx) : "no CUDA device (there is no CPU path)");
          if (ctx) rv_destroy(ctx);
          ctx = NULL;
          continue;
        }
      }
      std::vector<std::string> genes;
      for (auto& s : blk->specs) genes.push_back(s.gene);
      BatchTiming tm;
      int rc = somatic ? run_batch_somatic(ctx, c.P, batch, regs, genes, refseq, ref_lo, c.sample, chr, 3, L.halo, &blk->tsv, &tm, &blk->err)
                       : run_batch_simple(ctx, c.P, batch, regs, genes, refseq, ref_lo, c.sample, chr, 3, L.halo, &blk->tsv, &tm, &blk->err);
      if (rc != RV_OK) continue;
      blk->bases = tm.stats.n_aligned_bases;
      blk->reads = tm.stats.n_reads_kept;
      blk->lines = tm.n_lines;
      for (int k = 0; k < 2; ++k) { blk->cov_sum[k] = tm.cov_sum[k]; blk->cov_pos[k] = tm.cov_pos[k]; }
      blk->pileup_ms = tm.pileup_kernel_ms;
      blk->score_ms = tm.score_kernel_ms;
      if (tm.stats.n_unsupported)
        fprintf(stderr, "[warn] %lld reads hit a corner the device path refuses (see DESIGN.md)\n", (long long)tm.stats.n_unsupported);
    } catch (const std::exception& e) {
      blk->err = e.what();
    }
  }
  if (ctx) rv_destroy(ctx);
}

int main(int argc, char** argv) {
  double t0 = now_ms();
  Cli c;
  if (!parse(argc, argv, c)) return 1;
  rvio::BamReader hdr_reader;
  if (!hdr_reader.open(c.bam)) { fprintf(stderr, "cannot open %s\n", c.bam.c_str()); return 1; }
  if (c.sample.empty()) {  // SAMPLE_PATTERN fallbacks of Launcher.cpp:212-243 reduce to the file stem here
    size_t sl = c.bam.find_last_of('/');
    std::string base = sl == std::string::npos ? c.bam : c.bam.substr(sl + 1);
    c.sample = base.substr(0, base.find_first_of("._"));
  }
  std::vector<RegionSpec> specs;
  if (!build_regions(c, hdr_reader.header(), &specs) || specs.empty()) {
    fprintf(stderr, "no regions (give -R or -i)\n");
    return 1;
  }
  int ndev = rv_device_count();
  if (ndev <= 0) {
    fprintf(stderr, "rabbitvar_b200: no CUDA device visible; this build has no CPU path\n");
    return 3;
  }
  c.gpus = std::min(c.gpus, ndev);
  // blocks: consecutive regions of one contig, at most batch_regions each (a tile is the parity unit;
  // blocks never split a region)
  std::vector<Block> blocks;
  for (size_t i = 0; i < specs.size();) {
    Block b;
    size_t j = i;
    while (j < specs.size() && specs[j].chr == specs[i].chr && (int)(j - i) < c.batch_regions) b.specs.push_back(specs[j++]);
    blocks.push_back(b);
    i = j;
  }
  std::vector<std::vector<Block*> > per_gpu(c.gpus);
  for (size_t i = 0; i < blocks.size(); ++i) per_gpu[i * c.gpus / blocks.size()].push_back(&blocks[i]);
  std::vector<std::thread> th;
  for (int g = 0; g < c.gpus; ++g) th.emplace_back(worker, std::cref(c), g, per_gpu[g]);
  for (auto& t : th) t.join();
  FILE* out = fopen(c.out.c_str(), "wb");
  if (!out) { fprintf(stderr, "open file: %s error!\n", c.out.c_str()); return 1; }
  int64_t bases = 0, lines = 0;
  double kms = 0;
  int rc = 0;
  for (auto& b : blocks) {
    if (!b.err.empty()) { fprintf(stderr, "[error] %s\n", b.err.c_str()); rc = 2; }
    fwrite(b.tsv.data(), 1, b.tsv.size(), out);
    bases += b.bases;
    lines += b.lines;
    kms += b.pileup_ms + b.score_ms;
  }
  fclose(out);
  if (!c.bam2.empty()) {  // SomaticMode::process, somaticMode.cpp:916-929: average coverages (both over the tumor's sites)
    int64_t ts = 0, tp = 0, ns = 0;
    for (auto& b : blocks) { ts += b.cov_sum[0]; tp += b.cov_pos[0]; ns += b.cov_sum[1]; }
    FILE* info = fopen((c.out + ".info").c_str(), "wb");
    if (info) {
      const std::string s = std::to_string(ts / (double)tp) + "\n" + std::to_string(ns / (double)tp) + "\n";
      fwrite(s.data(), 1, s.size(), info);
      fclose(info);
    }
  }
  printf("[info] output file name: %s\n[info] regions: %zu blocks: %zu gpus: %d aligned bases: %lld variant lines: %lld kernel ms: %.3f\n",
         c.out.c_str(), specs.size(), blocks.size(), c.gpus, (long long)bases, (long long)lines, kms);
  printf("total time: %f s \n", (now_ms() - t0) / 1000.0);
  return rc;
}
