// file_pipeline.hpp — BAM/FASTA files in, TSV text out: the region loop of the reference
// (SimpleMode::process src/modes/simpleMode.cpp:210-387, SomaticMode::process src/modes/somaticMode.cpp:860-930)
// as a three-stage pipeline over JOBS (runs of consecutive regions of one contig):
//
//   decode threads  (--th of them)      one job each: BGZF inflate + BAM record parse of the job's spans for every
//                                       sample, FASTA slice, region descriptors.  Every thread owns its file handles
//                                       and its z_stream, as every OpenMP thread of the reference owns its BAM
//                                       handle (simpleMode.cpp:296-320); nothing is shared but the page cache.
//   GPU workers     (per device)        one rv_ctx (= one CUDA stream + device buffers) each: H2D, rv_pileup, host
//                                       hand-off (event reduce + realigner), rv_score, D2H, TSV assembly
//                                       (run_batch_simple / run_batch_somatic).  Workers of all devices pull from one
//                                       queue of decoded jobs, so the devices balance themselves.
//   writer                              the jobs' text in job order (replaces the `omp critical` write,
//                                       simpleMode.cpp:339-344; the reference prints in thread-completion order,
//                                       outputs are compared as sorted multisets — SURVEY Appendix A-17).
//
// A job's reads are fetched per CLUSTER of nearby regions (gap <= cluster_gap), not for the job's whole span: a sparse
// panel BED on a deep BAM does not pull the inter-region gaps into memory (the reference fetches per region).
#pragma once
#include "pipeline.hpp"
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

namespace rvhost {

struct FileRunConfig {
  std::string fasta, bam, bam2, sample;
  rv_params P;
  int ref_ext = 1200, nucl_ext = 0;
  int decode_threads = 1;
  int gpus = 1, first_device = 0;
  int workers_per_gpu = 3;
  int max_regions_per_job = 256;
  int64_t job_bytes = 24 << 20;  // compressed bytes (all samples) a job aims for
  int cluster_gap = 16384;  // (one index window: a scan starts at its window's first record anyway)
  int halo = 512;
  // decoded jobs waiting for a GPU worker are bounded by their bytes, not their number: while CUDA starts up (0.5-2 s
  // of a process's life) the decode threads should be able to run through a whole panel / exome sized input
  int64_t inflight_bytes = (int64_t)6 << 30;
  bool verbose = false;
  bool keep_contexts = false;  // do not rv_destroy the worker contexts at the end (a CLI about to exit)
  bool decode_only = false;  // measurement aid: run the decode stage alone (no device needed), print nothing
};

struct FileRunStats {
  int64_t bases = 0, reads = 0, lines = 0, n_jobs = 0, n_unsupported = 0, n_clipped = 0, dropped_keys = 0;
  int64_t cov_sum[2] = {0, 0}, cov_pos[2] = {0, 0};
  double pileup_kernel_ms = 0, score_kernel_ms = 0;
  double decode_thread_ms = 0, gpu_worker_ms = 0;  // summed over threads
  int64_t h2d_bytes = 0, d2h_bytes = 0, launches = 0;
  // gpu-worker time by stage, summed over the worker threads: context creation, read upload, pileup, event / row
  // fetch, host stage (reduce + realigner), patch upload, scoring, record fetch + TSV assembly
  double stage_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double first_job_ready_ms = 0, first_gpu_job_start_ms = 0, last_decode_done_ms = 0;  // since run_files began
};

struct FileJob {
  std::vector<RegionSpec> specs;  // same contig, in input order
  ReadBatch batch;
  std::vector<rv_region> regs;
  std::string refseq;
  int32_t ref_lo = 1, chr_len = 0;
  std::string tsv, err;
  BatchTiming tm;
  bool done = false;
  int64_t held_bytes = 0;  // decoded inputs this job holds until a GPU worker has taken them
  int64_t planned_bytes = 0;  // compressed bytes the planner counted for the job
};

// compressed offset at which the scan for `pos0` of `tid` starts (a monotone proxy for "where in the file")
inline uint64_t bai_coffset(const rvio::BaiIndex& bai, int tid, int64_t pos0) {
  uint64_t v = 0;
  if (!bai.start_offset(tid, pos0 < 0 ? 0 : pos0, &v)) return 0;
  return v >> 16;
}

// compressed bytes of the 16 kb index windows [w0, w1] of `tid`
inline int64_t window_bytes(const rvio::BaiIndex& bai, int tid, int64_t w0, int64_t w1) {
  if (w1 < w0) return 0;
  const uint64_t a = bai_coffset(bai, tid, w0 << 14), b = bai_coffset(bai, tid, (w1 + 1) << 14);
  return b > a ? (int64_t)(b - a) : 0;
}

// Cuts the region list into jobs: consecutive regions of one contig, at most max_regions_per_job, closed when the
// compressed bytes of the index windows the job's regions touch (each window once, summed over the samples) reach
// job_bytes.  Counting windows rather than region lengths keeps a panel of short amplicons on a deep BAM from
// collapsing into a few huge jobs: an amplicon's reads are all of its window's bytes.
inline void plan_jobs(const FileRunConfig& c, const std::vector<RegionSpec>& specs, const rvio::BamHeader& hdr,
                      const rvio::BaiIndex& bai, const rvio::BamHeader* hdr2, const rvio::BaiIndex* bai2,
                      std::vector<FileJob>* jobs) {
  jobs->clear();
  for (size_t i = 0; i < specs.size();) {
    FileJob j;
    const int tid = hdr.tid_of(specs[i].chr);
    const int tid2 = hdr2 ? hdr2->tid_of(specs[i].chr) : -1;
    int64_t bytes = 0, w_done = -1;  // windows up to w_done are counted (regions of a job mostly ascend)
    size_t k = i;
    while (k < specs.size() && specs[k].chr == specs[i].chr && (int)(k - i) < c.max_regions_per_job) {
      if (k > i && bytes >= c.job_bytes) break;
      int64_t w0 = ((int64_t)specs[k].start - 1) >> 14, w1 = (int64_t)specs[k].end >> 14;
      if (w0 < 0) w0 = 0;
      if (w1 < w_done - 64 || w0 > w_done) w_done = w0 - 1;  // a jump backwards / a gap: start counting afresh
      if (w0 <= w_done) w0 = w_done + 1;
      if (tid >= 0) bytes += window_bytes(bai, tid, w0, w1);
      if (bai2 && tid2 >= 0) bytes += window_bytes(*bai2, tid2, w0, w1);
      if (w1 > w_done) w_done = w1;
      j.specs.push_back(specs[k++]);
    }
    j.planned_bytes = bytes;
    jobs->push_back(std::move(j));
    i = k;
  }
}

struct SampleFiles {
  rvio::SpanScanner scan;
  const rvio::BaiIndex* bai = NULL;
};

// Decode of one job: reads of every cluster of regions (per sample), region descriptors, reference slice.
inline void decode_job(const FileRunConfig& c, SampleFiles* samples, int n_samples, rvio::Fasta& fa, FileJob* job) {
  const std::string& chr = job->specs[0].chr;
  const rvio::BamHeader& hdr = samples[0].scan.header();
  const int tid = hdr.tid_of(chr);
  if (tid < 0) { job->err = "contig not in BAM: " + chr; return; }
  job->chr_len = hdr.lens[(size_t)tid];
  // clusters of regions whose reads are fetched together: regions in ascending order, gaps <= cluster_gap
  std::vector<size_t> order(job->specs.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return job->specs[a].start < job->specs[b].start; });
  struct Cluster { int32_t lo, hi; std::vector<size_t> members; };
  std::vector<Cluster> clusters;
  for (size_t oi = 0; oi < order.size(); ++oi) {
    const RegionSpec& s = job->specs[order[oi]];
    if (clusters.empty() || s.start > clusters.back().hi + c.cluster_gap) {
      Cluster cl;
      cl.lo = s.start; cl.hi = s.end;
      clusters.push_back(cl);
    }
    clusters.back().hi = std::max(clusters.back().hi, s.end);
    clusters.back().members.push_back(order[oi]);
  }
  job->batch.clear();
  {  // room for the job's reads up front: deflate shrinks BAM payloads ~3.5x, a record is ~300 bytes
    int64_t cbytes = 0;
    for (int s = 0; s < n_samples; ++s) {
      const int tid_s = samples[s].scan.header().tid_of(chr);
      if (tid_s < 0) continue;
      for (size_t ci = 0; ci < clusters.size(); ++ci) {
        const uint64_t a = bai_coffset(*samples[s].bai, tid_s, (int64_t)clusters[ci].lo - 1), b = bai_coffset(*samples[s].bai, tid_s, (int64_t)clusters[ci].hi + (1 << 14));
        if (b > a) cbytes += (int64_t)(b - a);
      }
    }
    cbytes = std::min<int64_t>(cbytes, (int64_t)1 << 30);
    job->batch.pool.reserve((size_t)(cbytes * 4));
    job->batch.reads.reserve((size_t)(cbytes * 4 / 250));
  }
  // one allocation for the pool instead of a doubling series (untouched reserve costs address space only): a level-1
  // BAM inflates ~3.5x, and cigar | bases | qualities are ~0.8 of a record
  job->batch.pool.reserve((size_t)std::min<int64_t>(job->planned_bytes * 4 + (1 << 20), (int64_t)3 << 30));
  job->batch.reads.reserve((size_t)std::min<int64_t>(job->planned_bytes / 40 + 1024, (int64_t)1 << 24));
  job->regs.assign(job->specs.size() * (size_t)n_samples, rv_region());
  int32_t smin = job->specs[0].start, smax = job->specs[0].end;
  for (size_t i = 0; i < job->specs.size(); ++i) { smin = std::min(smin, job->specs[i].start); smax = std::max(smax, job->specs[i].end); }
  for (int s = 0; s < n_samples; ++s) {
    const int tid_s = samples[s].scan.header().tid_of(chr);
    if (tid_s < 0) { job->err = "contig not in the second BAM: " + chr; return; }
    for (size_t ci = 0; ci < clusters.size(); ++ci) {
      const int64_t r0 = (int64_t)job->batch.reads.size();
      // (the margin: reads that overlap no region of the span but pass an indel found just outside one — the realigner's
      //  noPassingReads looks them up, VariationRealigner.cpp:1447-1490; no region's read slice includes them)
      load_span_fast(samples[s].scan, *samples[s].bai, tid_s, std::max(1, clusters[ci].lo - c.halo),
                     std::min(job->chr_len, clusters[ci].hi + c.halo), &job->batch);
      const int64_t r1 = (int64_t)job->batch.reads.size();
      std::vector<RegionSpec> sub;
      for (size_t m = 0; m < clusters[ci].members.size(); ++m) sub.push_back(job->specs[clusters[ci].members[m]]);
      std::vector<rv_region> rr;
      make_regions_range(job->batch, sub, job->chr_len, c.ref_ext, c.nucl_ext, r0, r1, &rr);
      for (size_t m = 0; m < clusters[ci].members.size(); ++m)
        job->regs[(size_t)s * job->specs.size() + clusters[ci].members[m]] = rr[m];
    }
  }
  job->ref_lo = std::max(1, smin - c.ref_ext - c.nucl_ext - 100);
  const int32_t ref_hi = std::min(job->chr_len, smax + c.ref_ext + c.nucl_ext + 100);
  if (!fa.fetch(chr, job->ref_lo, ref_hi, &job->refseq) || (int64_t)job->refseq.size() != (int64_t)ref_hi - job->ref_lo + 1) {
    job->err = "cannot read " + chr + ":" + std::to_string(job->ref_lo) + "-" + std::to_string(ref_hi) + " from the FASTA (contig missing from the .fai, or file truncated)";
    return;
  }
  for (size_t i = 0; i < job->refseq.size(); ++i) job->refseq[i] = (char)toupper((unsigned char)job->refseq[i]);
}

// capacity a job needs from its context
inline void job_limits(const FileJob& j, int halo, bool somatic, rv_limits* need) {
  rv_default_limits(need);
  int64_t npos = 0;
  for (size_t r = 0; r < j.regs.size(); ++r) npos += j.regs[r].end - j.regs[r].start + 1 + 2 * halo;
  need->halo = halo;
  need->max_reads = (int64_t)j.batch.reads.size() + 1024;
  need->max_read_bytes = (int64_t)j.batch.pool.size() + 4096;
  need->max_positions = npos + 1024;
  need->max_regions = (int32_t)j.regs.size() + 8;
  need->max_events = std::max<int64_t>(1 << 17, (int64_t)j.batch.reads.size());
  need->max_variants = (somatic ? 3 : 1) * npos + 1024;
  need->max_patch = std::max<int64_t>(1 << 16, (int64_t)j.batch.reads.size() / 2);
  need->max_ref_bases = (int64_t)j.refseq.size() + 1024;
}
inline bool limits_fit(const rv_limits& have, const rv_limits& need) {
  return have.max_reads >= need.max_reads && have.max_read_bytes >= need.max_read_bytes &&
         have.max_positions >= need.max_positions && have.max_regions >= need.max_regions &&
         have.max_events >= need.max_events && have.max_variants >= need.max_variants &&
         have.max_patch >= need.max_patch && have.max_ref_bases >= need.max_ref_bases && have.halo == need.halo &&
         have.max_sparse_obs >= need.max_sparse_obs;
}

// Runs every region of `specs` (input order is output order).  Returns 0, or 2 when a job failed (its regions print
// nothing, `errors` says why), or 3 without a CUDA device.
inline int run_files(const FileRunConfig& c, const std::vector<RegionSpec>& specs, std::string* tsv, FileRunStats* st,
                     std::vector<std::string>* errors) {
  const bool somatic = !c.bam2.empty();
  const int n_samples = somatic ? 2 : 1;
  rvio::BaiIndex bai, bai2;
  rvio::BamReader hdr_reader, hdr_reader2;
  if (!hdr_reader.open(c.bam)) { errors->push_back("cannot open " + c.bam); return 1; }
  if (!bai.load(c.bam + ".bai")) { errors->push_back("cannot open " + c.bam + ".bai"); return 1; }
  if (somatic) {
    if (!hdr_reader2.open(c.bam2)) { errors->push_back("cannot open " + c.bam2); return 1; }
    if (!bai2.load(c.bam2 + ".bai")) { errors->push_back("cannot open " + c.bam2 + ".bai"); return 1; }
  }
  std::vector<FileJob> jobs;
  plan_jobs(c, specs, hdr_reader.header(), bai, somatic ? &hdr_reader2.header() : NULL, somatic ? &bai2 : NULL, &jobs);
  const int n_jobs = (int)jobs.size();
  const int n_gpu_workers = std::max(1, c.gpus * c.workers_per_gpu);
  const int n_decode = std::max(1, std::min(c.decode_threads, n_jobs));
  // decoded-but-unprocessed jobs are bounded (memory): a decode thread waits for a slot once more than min_inflight
  // jobs are out and the decoded bytes held exceed c.inflight_bytes
  const int min_inflight = n_decode + 2 * n_gpu_workers;
  int64_t held_bytes = 0;
  std::mutex mu;
  std::condition_variable cv_ready, cv_slot;
  std::deque<int> ready;
  int inflight = 0, next_job = 0, decoders_left = n_decode;
  std::atomic<long long> dec_us(0), gpu_us(0), launches(0), create_us(0);
  const double t_begin = now_ms();
  std::atomic<long long> first_ready_us(-1), first_gpu_us(-1), last_dec_us(0);
  std::vector<std::thread> threads;
  for (int d = 0; d < n_decode; ++d)
    threads.emplace_back([&, d]() {
      (void)d;
      SampleFiles sf[2];
      rvio::Fasta fa;
      bool ok = sf[0].scan.open(c.bam) && fa.open(c.fasta);
      sf[0].bai = &bai;
      if (ok && somatic) { ok = sf[1].scan.open(c.bam2); sf[1].bai = &bai2; }
      for (;;) {
        int ji;
        {
          std::unique_lock<std::mutex> lk(mu);
          cv_slot.wait(lk, [&]() { return inflight < min_inflight || held_bytes < c.inflight_bytes || next_job >= n_jobs; });
          if (next_job >= n_jobs) break;
          ji = next_job++;
          inflight++;
        }
        const double t0 = now_ms();
        FileJob& job = jobs[(size_t)ji];
        if (!ok) job.err = "cannot open BAM/FASTA (" + c.bam + ", " + c.fasta + ")";
        else {
          try { decode_job(c, sf, n_samples, fa, &job); }
          catch (const std::exception& e) { job.err = e.what(); }
        }
        dec_us += (long long)((now_ms() - t0) * 1000.0);
        {
          const long long t_us = (long long)((now_ms() - t_begin) * 1000.0);
          long long none = -1;
          first_ready_us.compare_exchange_strong(none, t_us);
          long long prev = last_dec_us.load();
          while (prev < t_us && !last_dec_us.compare_exchange_weak(prev, t_us)) {}
        }
        job.held_bytes = (int64_t)(job.batch.pool.size() + job.batch.reads.size() * sizeof(rv_read) + job.refseq.size());
        {
          std::lock_guard<std::mutex> lk(mu);
          held_bytes += job.held_bytes;
          ready.push_back(ji);
        }
        cv_ready.notify_one();
      }
      {
        std::lock_guard<std::mutex> lk(mu);
        decoders_left--;
      }
      cv_ready.notify_all();
    });
  const int host_thr = std::max(1, c.decode_threads / n_gpu_workers);
  for (int w = 0; w < n_gpu_workers; ++w)
    threads.emplace_back([&, w]() {
      host_threads_override() = host_thr;
      int n_dev = rv_device_count();
      if (n_dev <= 0) n_dev = 1;
      const int device = (c.first_device + w % std::max(1, c.gpus)) % n_dev;
      if (!c.decode_only) rv_warmup(device);  // (waits for the CUDA start-up the caller began on another thread)
      rv_ctx* ctx = NULL;
      rv_limits have;
      rv_default_limits(&have);
      for (;;) {
        int ji;
        {
          std::unique_lock<std::mutex> lk(mu);
          cv_ready.wait(lk, [&]() { return !ready.empty() || decoders_left == 0; });
          if (ready.empty()) break;
          ji = ready.front();
          ready.pop_front();
        }
        const double t0 = now_ms();
        {
          long long none = -1;
          first_gpu_us.compare_exchange_strong(none, (long long)((t0 - t_begin) * 1000.0));
        }
        FileJob& job = jobs[(size_t)ji];
        if (job.err.empty() && c.decode_only) {
          job.tm = BatchTiming();
          memset(&job.tm, 0, sizeof job.tm);
          job.tm.stats.n_reads_kept = (int64_t)job.batch.reads.size();
        } else if (job.err.empty()) {
          try {
            rv_limits need;
            job_limits(job, c.halo, somatic, &need);
            // a job that overflows the event / variant buffers is re-run once with larger ones
            for (int attempt = 0; attempt < 3; ++attempt) {
              if (!ctx || !limits_fit(have, need)) {
                if (ctx) { launches += rv_launch_count(ctx); rv_destroy(ctx); ctx = NULL; }
                rv_limits L = need;  // head-room so that the following jobs fit as well
                L.max_reads += L.max_reads / 4; L.max_read_bytes += L.max_read_bytes / 4; L.max_positions += L.max_positions / 4;
                L.max_events += L.max_events / 4; L.max_patch += L.max_patch / 4; L.max_ref_bases += L.max_ref_bases / 4;
                L.max_variants = (somatic ? 3 : 1) * L.max_positions + 1024;
                L.max_regions = std::max(L.max_regions, c.max_regions_per_job * n_samples + 8);
                if (L.max_read_bytes >= ((int64_t)1 << 32) - 4096) L.max_read_bytes = need.max_read_bytes;
                const double tc = now_ms();
                const int rc = rv_create(&ctx, device, &c.P, &L);
                create_us += (long long)((now_ms() - tc) * 1000.0);
                if (rc != RV_OK) {
                  job.err = std::string("rv_create: ") + (ctx ? rv_last_error(ctx) : "no CUDA device (there is no CPU path)");
                  if (ctx) rv_destroy(ctx);
                  ctx = NULL;
                  break;
                }
                have = L;
              }
              std::vector<std::string> genes;
              for (size_t i = 0; i < job.specs.size(); ++i) genes.push_back(job.specs[i].gene);
              job.tsv.clear();
              job.err.clear();
              const int rc = somatic ? run_batch_somatic(ctx, c.P, job.batch, job.regs, genes, job.refseq, job.ref_lo, c.sample, job.specs[0].chr, 3, c.halo, &job.tsv, &job.tm, &job.err)
                                     : run_batch_simple(ctx, c.P, job.batch, job.regs, genes, job.refseq, job.ref_lo, c.sample, job.specs[0].chr, 3, c.halo, &job.tsv, &job.tm, &job.err);
              if (rc == RV_OK) { job.err.clear(); break; }
              if (rc != RV_ERR_OVERFLOW) break;
              need.max_events *= 4; need.max_variants *= 2; need.max_patch *= 4;
              need.max_sparse_obs = std::max<int64_t>(4 << 20, 8 * need.max_reads) << (4 * (attempt + 1));
            }
          } catch (const std::exception& e) {
            job.err = e.what();
          }
        }
        // the job's inputs are not needed any more
        ReadBatch().swap(job.batch);
        std::string().swap(job.refseq);
        job.done = true;
        gpu_us += (long long)((now_ms() - t0) * 1000.0);
        {
          std::lock_guard<std::mutex> lk(mu);
          inflight--;
          held_bytes -= job.held_bytes;
        }
        cv_slot.notify_all();
      }
      if (ctx) {
        launches += rv_launch_count(ctx);
        if (!c.keep_contexts) rv_destroy(ctx);
      }
    });
  for (size_t i = 0; i < threads.size(); ++i) threads[i].join();
  int rc = 0;
  size_t total = 0;
  for (int j = 0; j < n_jobs; ++j) total += jobs[(size_t)j].tsv.size();
  tsv->clear();
  tsv->reserve(total);
  for (int j = 0; j < n_jobs; ++j) {
    FileJob& job = jobs[(size_t)j];
    if (!job.err.empty()) {
      errors->push_back(job.specs[0].chr + ":" + std::to_string(job.specs[0].start) + " (" + std::to_string(job.specs.size()) + " regions): " + job.err);
      rc = 2;
      continue;
    }
    tsv->append(job.tsv);
    st->bases += job.tm.stats.n_aligned_bases;
    st->reads += job.tm.stats.n_reads_kept;
    st->lines += job.tm.n_lines;
    st->n_unsupported += job.tm.stats.n_unsupported;
    st->n_clipped += job.tm.stats.n_clipped;
    st->pileup_kernel_ms += job.tm.pileup_kernel_ms;
    st->score_kernel_ms += job.tm.score_kernel_ms;
    st->h2d_bytes += job.tm.h2d_bytes;
    st->d2h_bytes += job.tm.d2h_bytes;
    st->stage_ms[1] += job.tm.push_ms; st->stage_ms[2] += job.tm.pileup_ms; st->stage_ms[3] += job.tm.fetch_ms;
    st->stage_ms[4] += job.tm.host_ms; st->stage_ms[5] += job.tm.patch_ms; st->stage_ms[6] += job.tm.score_ms;
    st->stage_ms[7] += job.tm.assemble_ms;
    for (int k = 0; k < 2; ++k) { st->cov_sum[k] += job.tm.cov_sum[k]; st->cov_pos[k] += job.tm.cov_pos[k]; }
  }
  st->n_jobs = n_jobs;
  st->decode_thread_ms = dec_us.load() / 1000.0;
  st->gpu_worker_ms = gpu_us.load() / 1000.0;
  st->launches = launches.load();
  st->stage_ms[0] = create_us.load() / 1000.0;
  st->first_job_ready_ms = first_ready_us.load() / 1000.0;
  st->first_gpu_job_start_ms = first_gpu_us.load() / 1000.0;
  st->last_decode_done_ms = last_dec_us.load() / 1000.0;
  st->dropped_keys = dropped_patch_keys().load();
  return rc;
}

}  // namespace rvhost
