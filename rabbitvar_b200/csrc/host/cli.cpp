// cli.cpp — command-line front end that keeps the reference's CLI, region/BED handling and TSV output
// for the pileup-and-score path (reference src/Launcher.cpp:295-496 cmdParse, src/RegionBuilder.cpp:45-109
// buildRegions, :179-200 buildRegionFromConfiguration, src/modes/simpleMode.cpp:210-387 SimpleMode::process).
//
// The OpenMP region loop of the reference becomes the three-stage job pipeline of file_pipeline.hpp: --th decode
// threads (BGZF inflate + BAM parse, one file handle each) feed GPU worker contexts on --gpus devices, the host
// concatenates the jobs' TSV in region order (replaces the `omp critical` write, simpleMode.cpp:339).
// No collective sits on the path.
//
// Supported: simple mode (-b one BAM) and paired somatic mode (-b 'tumor.bam|normal.bam', somaticMode.cpp), -R or
// -i BED with -c/-S/-E/-g, the scoring/filter flags below.
#include "../../../include/rabbitvar_b200.h"
#include "file_pipeline.hpp"
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
  int nucl_ext = 0, ref_ext = 1200, gpus = 1, threads = 1, batch_regions = 256, halo = 512, workers = 4, job_mb = 12, device = 0, inflight_gb = 6;
  bool auto_resize = false, threads_set = false, decode_only = false;
  rv_params P;
};

static void usage() {
  fprintf(stderr,
          "usage: rabbitvar_b200 -G ref.fa -b in.bam (-R chr:start-end | -i regions.bed -c 1 -S 2 -E 3 -g 4) [options]\n"
          "  -N name  -f freq  -k 0|1  -3  -u  --UN  -p  -t  --fisher  -q phred  -m mismatches  -X ext  -P pos  -r minr\n"
          "  -B minbias  -Q mapq  -o qratio  -O mapq  -V lofreq  -M minmatch  -T trim  -F hexfilter  -x extend  -Y refext\n"
          "  -z  --auto_resize  --th n  --gpus n  --device first  --workers n (GPU contexts per device)  --halo n  --job-mb n  --inflight-gb n (decoded input held ahead of the GPU)  --out file\n");
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
    else if (a == "--th") { c.threads = std::max(1, atoi(val().c_str())); c.threads_set = true; }
    else if (a == "--halo") c.halo = std::max(64, atoi(val().c_str()));
    else if (a == "--workers") c.workers = std::max(1, atoi(val().c_str()));
    else if (a == "--device") c.device = std::max(0, atoi(val().c_str()));
    else if (a == "--job-mb") c.job_mb = std::max(1, atoi(val().c_str()));
    else if (a == "--inflight-gb") c.inflight_gb = std::max(0, atoi(val().c_str()));
    else if (a == "--auto_resize") c.auto_resize = true;
    else if (a == "--decode-only") c.decode_only = true;
    else if (a == "--gpus") c.gpus = std::max(1, atoi(val().c_str()));
    else if (a == "--batch-regions") c.batch_regions = std::max(1, atoi(val().c_str()));
    else if (a == "-y" || a == "--verbose" || a == "--chimeric" || a == "--deldupvar") {}
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
    // --auto_resize: RegionBuilder::AdjustRegionSize (RegionBuilder.cpp:16-38, REGION_SIZE_MAX 10000) — literal: the
    // pieces are [s, s+10000], [s+10000, s+20000], ...: neighbouring pieces share their boundary position
    if (c.auto_resize && r.end - r.start > 10000) {
      int s0 = r.start;
      while (r.end - s0 > 10000) {
        RegionSpec piece = r;
        piece.start = s0;
        piece.end = s0 + 10000;
        out->push_back(piece);
        s0 += 10000;
      }
      r.start = s0;
    }
    out->push_back(r);
  }
  return true;
}

int main(int argc, char** argv) {
  double t0 = now_ms();
  Cli c;
  if (!parse(argc, argv, c)) return 1;
  // CUDA start-up (driver initialisation, primary context, module load) runs beside the header / index / BED parsing.
  // With one device in use the others are hidden from the driver first: initialising eight GPUs costs several times
  // what initialising one does.
  // (--device counts within CUDA_VISIBLE_DEVICES when a launcher has set one — torchrun jobs usually run with all eight
  // listed —, so the entries [device, device + gpus) of that list are the ones kept)
  if (c.gpus >= 1 && c.device >= 0) {
    const char* vis = getenv("CUDA_VISIBLE_DEVICES");
    std::vector<std::string> ids;
    if (vis == NULL) {
      for (int k = 0; k < c.device + c.gpus; ++k) ids.push_back(std::to_string(k));
    } else {
      std::string cur;
      for (const char* p = vis;; ++p) {
        if (*p == ',' || *p == 0) { ids.push_back(cur); cur.clear(); if (*p == 0) break; }
        else if (*p != ' ') cur += *p;
      }
    }
    std::string keep;
    for (size_t k = (size_t)c.device; k < ids.size() && k < (size_t)(c.device + c.gpus); ++k) {
      if (ids[k].empty()) { keep.clear(); break; }
      keep += (keep.empty() ? "" : ",") + ids[k];
    }
    if (!keep.empty()) {
      setenv("CUDA_VISIBLE_DEVICES", keep.c_str(), 1);
      c.device = 0;
    }
  }
  double cuda_init_ms = 0;
  std::thread cuda_init;
  if (!c.decode_only) cuda_init = std::thread([&]() { const double a = now_ms(); rv_warmup(c.device); cuda_init_ms = now_ms() - a; });
  struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{cuda_init};
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
  const double t_parsed = now_ms();
  // (the decode threads start right away; the GPU workers wait for the CUDA start-up inside rv_create)
  FileRunConfig fc;
  fc.fasta = c.fasta; fc.bam = c.bam; fc.bam2 = c.bam2; fc.sample = c.sample;
  fc.P = c.P;
  fc.ref_ext = c.ref_ext; fc.nucl_ext = c.nucl_ext;
  fc.decode_threads = c.threads;                 // --th, default 1 like the reference (Launcher.cpp:474)
  fc.first_device = c.device;
  fc.gpus = std::max(1, c.gpus);
  fc.decode_only = c.decode_only;
  fc.workers_per_gpu = c.workers;
  fc.max_regions_per_job = c.batch_regions;
  fc.job_bytes = (int64_t)c.job_mb << 20;
  fc.inflight_bytes = (int64_t)c.inflight_gb << 30;  // decoded input the decode threads may run ahead of the GPU workers
  fc.halo = c.halo;
  fc.keep_contexts = true;  // the process exits right after the run
  std::string tsv;
  FileRunStats st;
  std::vector<std::string> errors;
  int rc = run_files(fc, specs, &tsv, &st, &errors);
  const double t_ran = now_ms();
  if (cuda_init.joinable()) cuda_init.join();
  if (!c.decode_only && rv_device_count() <= 0) {
    fprintf(stderr, "rabbitvar_b200: no CUDA device visible; this build has no CPU path\n");
    return 3;
  }
  for (size_t i = 0; i < errors.size(); ++i) fprintf(stderr, "[error] %s\n", errors[i].c_str());
  if (rc == 1) return 1;
  FILE* out = fopen(c.out.c_str(), "wb");
  if (!out) { fprintf(stderr, "open file: %s error!\n", c.out.c_str()); return 1; }
  fwrite(tsv.data(), 1, tsv.size(), out);
  fclose(out);
  if (!c.bam2.empty()) {  // SomaticMode::process, somaticMode.cpp:916-929: average coverages (both over the tumor's sites)
    FILE* info = fopen((c.out + ".info").c_str(), "wb");
    if (info) {
      const std::string s = std::to_string(st.cov_sum[0] / (double)st.cov_pos[0]) + "\n" + std::to_string(st.cov_sum[1] / (double)st.cov_pos[0]) + "\n";
      fwrite(s.data(), 1, s.size(), info);
      fclose(info);
    }
  }
  if (st.n_unsupported)
    fprintf(stderr, "[warn] %lld reads hit a corner the device path refuses (see DESIGN.md): their observations are missing\n", (long long)st.n_unsupported);
  if (st.dropped_keys)
    fprintf(stderr, "[warn] %lld allele keys longer than %d characters were dropped by the host stage\n", (long long)st.dropped_keys, RV_PATCH_KEY_MAX);
  if (st.n_clipped)
    fprintf(stderr, "[warn] %lld observations fell outside the table halo of %d positions and were dropped (--halo)\n", (long long)st.n_clipped, c.halo);
  printf("[info] output file name: %s\n[info] regions: %zu jobs: %lld gpus: %d x %d contexts, decode threads: %d, aligned bases: %lld variant lines: %lld "
         "kernel ms: %.3f, decode thread-ms %.0f, gpu-worker thread-ms %.0f, launches %lld, h2d bytes %lld, d2h bytes %lld\n",
         c.out.c_str(), specs.size(), (long long)st.n_jobs, fc.gpus, fc.workers_per_gpu, fc.decode_threads, (long long)st.bases, (long long)st.lines,
         st.pileup_kernel_ms + st.score_kernel_ms, st.decode_thread_ms, st.gpu_worker_ms, (long long)st.launches,
         (long long)st.h2d_bytes, (long long)st.d2h_bytes);
  printf("[info] gpu-worker thread-ms by stage: contexts %.0f, upload %.0f, pileup %.0f, fetch %.0f, host stage %.0f, patch %.0f, score %.0f, "
         "records + TSV %.0f; first job decoded at %.0f ms, last at %.0f ms, first job taken by a GPU worker at %.0f ms\n",
         st.stage_ms[0], st.stage_ms[1], st.stage_ms[2], st.stage_ms[3], st.stage_ms[4], st.stage_ms[5], st.stage_ms[6], st.stage_ms[7],
         st.first_job_ready_ms, st.last_decode_done_ms, st.first_gpu_job_start_ms);
  printf("[info] CUDA_VISIBLE_DEVICES of this run: %s\n", getenv("CUDA_VISIBLE_DEVICES") ? getenv("CUDA_VISIBLE_DEVICES") : "(unset)");
  printf("[info] timeline ms: inputs parsed %.0f, cuda start-up %.0f (beside the decode threads), pipeline done %.0f, output written %.0f\n",
         t_parsed - t0, cuda_init_ms, t_ran - t0, now_ms() - t0);
  printf("total time: %f s \n", (now_ms() - t0) / 1000.0);
  // the output is on disk: leave without tearing the CUDA context down (hundreds of milliseconds of cudaFree)
  fflush(stdout);
  fflush(stderr);
  _exit(rc);
}
