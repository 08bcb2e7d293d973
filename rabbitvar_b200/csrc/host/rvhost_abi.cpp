// rvhost_abi.cpp — C ABI over the host side of the path (include/rabbitvar_b200_host.h).
#include "../../../include/rabbitvar_b200_host.h"
#include "file_pipeline.hpp"
#include <array>
#include <atomic>
#include <map>
#include <mutex>
#include <thread>

using namespace rvhost;

#include <cuda_runtime_api.h>

struct rvh_batch {
  ReadBatch b;
  bool pinned;
  rvh_batch() : pinned(false) {}
  ~rvh_batch() {
    if (pinned) {
      cudaHostUnregister((void*)b.reads.data());
      cudaHostUnregister((void*)b.pool.data());
    }
  }
};

static thread_local std::string g_err;
static std::mutex g_tsv_mu;
static std::map<rv_ctx*, std::string> g_tsv;  // per-context output buffers

// ---- pipelined region loop --------------------------------------------------------------------------
struct rvh_pipeline {
  int device;
  int n_workers;
  std::vector<rv_ctx*> ctx;
  std::vector<rv_limits> lim;        // capacity the worker contexts were created with
  std::vector<const char*> ref_seen; // reference slice resident on each worker context
  std::vector<int32_t> ref_lo_seen;
  std::vector<int64_t> ref_n_seen;
  std::string tsv;
  std::string refseq;
  int64_t launches_retired;
};

static bool limits_cover(const rv_limits& have, const rv_limits& need) {
  return have.max_reads >= need.max_reads && have.max_read_bytes >= need.max_read_bytes &&
         have.max_positions >= need.max_positions && have.max_regions >= need.max_regions &&
         have.max_events >= need.max_events && have.max_variants >= need.max_variants &&
         have.max_patch >= need.max_patch && have.max_ref_bases >= need.max_ref_bases && have.halo == need.halo &&
         have.max_sparse_obs >= need.max_sparse_obs;
}

extern "C" {

rvh_pipeline* rvh_pipeline_create(int device, int n_workers) {
  if (n_workers < 1) n_workers = 1;
  if (n_workers > 64) n_workers = 64;
  rvh_pipeline* p = new rvh_pipeline();
  p->device = device;
  p->n_workers = n_workers;
  p->ctx.assign((size_t)n_workers, (rv_ctx*)NULL);
  p->lim.resize((size_t)n_workers);
  p->ref_seen.assign((size_t)n_workers, (const char*)NULL);
  p->ref_lo_seen.assign((size_t)n_workers, 0);
  p->ref_n_seen.assign((size_t)n_workers, 0);
  p->launches_retired = 0;
  return p;
}

void rvh_pipeline_destroy(rvh_pipeline* p) {
  if (!p) return;
  for (size_t i = 0; i < p->ctx.size(); ++i) rv_destroy(p->ctx[i]);
  delete p;
}

int64_t rvh_pipeline_launch_count(const rvh_pipeline* p) {
  if (!p) return 0;
  int64_t n = p->launches_retired;
  for (size_t i = 0; i < p->ctx.size(); ++i) n += rv_launch_count(p->ctx[i]);
  return n;
}

static int pipeline_run(rvh_pipeline* p, const rv_params* params, const rvh_batch* batch, const rv_region* regions_in,
                        int32_t n_regions_in, int32_t chunk_regions, const char* ref_bases, int32_t ref_lo, int64_t ref_n,
                        const char* sample, const char* chr, const char** tsv_out, int64_t* tsv_len, rvh_timing* timing,
                        bool paired) {
  if (!p || !params || !batch || (!regions_in && n_regions_in) || !ref_bases || !tsv_out || !tsv_len || n_regions_in < 0)
    return RV_ERR_ARG;
  if (paired && (n_regions_in % 2)) return RV_ERR_ARG;
  const double t_enter = now_ms();
  try {
    // paired: regions_in = n tumor tiles followed by the same n tiles of the normal sample; a chunk takes tiles
    // [r0, r1) of both.  The loops below run over tiles.
    const rv_region* regions = regions_in;
    const int32_t n_regions = paired ? n_regions_in / 2 : n_regions_in;
    const int per_tile = paired ? 2 : 1;
    if (chunk_regions < 1) chunk_regions = 1;
    // chunk boundaries (in tiles): full-size chunks first, shrinking ones over the last stretch so that the workers
    // finish together (a worker runs its chunk's stages back to back; the tail would otherwise idle the other workers
    // for up to one chunk's duration)
    std::vector<int> cb(1, 0);
    while (cb.back() < n_regions) {
      const int left = n_regions - cb.back();
      int size = chunk_regions;
      // ... and growing ones at the start: every worker begins with an upload, small first chunks put the first
      // kernels and host stages behind a short copy instead of all workers' full-size ones
      const int wave = p->n_workers > 1 ? (int)(cb.size() - 1) / p->n_workers : 2;
      if (wave == 0) size = std::max(1, chunk_regions / 4);
      else if (wave == 1) size = std::max(1, chunk_regions / 2);
      if (p->n_workers > 1 && left < 2 * p->n_workers * chunk_regions)
        size = std::max(1, std::min(size, (left + 2 * p->n_workers - 1) / (2 * p->n_workers)));
      cb.push_back(cb.back() + std::min(size, left));
    }
    const int n_chunks = (int)cb.size() - 1;
    const ReadBatch& B = batch->b;
    const int halo = 512;
    // capacity one chunk needs
    rv_limits need;
    rv_default_limits(&need);
    need.halo = halo;
    need.max_reads = 1024; need.max_read_bytes = 4096; need.max_positions = 1024;
    need.max_regions = per_tile * chunk_regions + 8;
    // read range of every chunk, per sample: [c][2*k] = lo, [c][2*k+1] = hi
    std::vector<std::array<int64_t, 4> > c_rng((size_t)n_chunks);
    for (int c = 0; c < n_chunks; ++c) {
      const int r0 = cb[(size_t)c], r1 = cb[(size_t)c + 1];
      int64_t npos = 0, reads = 0, bytes = 0;
      for (int k = 0; k < per_tile; ++k) {
        int64_t lo = -1, hi = -1;
        for (int r = r0; r < r1; ++r) {
          const rv_region& g = regions[r + k * n_regions];
          npos += g.end - g.start + 1 + 2 * halo;
          if (g.read_hi <= g.read_lo) continue;
          if (lo < 0 || g.read_lo < lo) lo = g.read_lo;
          if (g.read_hi > hi) hi = g.read_hi;
        }
        if (lo < 0) lo = hi = 0;
        c_rng[(size_t)c][2 * k] = lo;
        c_rng[(size_t)c][2 * k + 1] = hi;
        if (hi > lo) {
          const int64_t p_lo = (int64_t)B.reads[(size_t)lo].data_off16 * 16;
          const int64_t p_hi = hi < (int64_t)B.reads.size() ? (int64_t)B.reads[(size_t)hi].data_off16 * 16 : (int64_t)B.pool.size();
          bytes += p_hi - p_lo + 16;
          reads += hi - lo;
        }
      }
      if (!paired) { c_rng[(size_t)c][2] = c_rng[(size_t)c][3] = 0; }
      need.max_reads = std::max<int64_t>(need.max_reads, reads + 64);
      need.max_read_bytes = std::max<int64_t>(need.max_read_bytes, bytes + 256);
      need.max_positions = std::max<int64_t>(need.max_positions, npos + 64);
    }
    need.max_events = std::max<int64_t>(1 << 18, need.max_reads);
    need.max_variants = 3 * need.max_positions + 1024;
    need.max_patch = std::max<int64_t>(1 << 18, need.max_reads / 2);
    need.max_ref_bases = ref_n + 64;
    for (int w = 0; w < p->n_workers; ++w) {
      if (p->ctx[(size_t)w] && limits_cover(p->lim[(size_t)w], need)) continue;
      if (p->ctx[(size_t)w]) {
        p->launches_retired += rv_launch_count(p->ctx[(size_t)w]);
        rv_destroy(p->ctx[(size_t)w]);
        p->ctx[(size_t)w] = NULL;
      }
      rv_limits L = need;  // head-room so that similar batches do not re-allocate
      L.max_reads += L.max_reads / 8; L.max_read_bytes += L.max_read_bytes / 8; L.max_positions += L.max_positions / 8;
      L.max_variants = 3 * L.max_positions + 1024;
      int rc = rv_create(&p->ctx[(size_t)w], p->device, params, &L);
      if (rc != RV_OK) {
        g_err = std::string("rv_create (pipeline worker): ") + rv_last_error(p->ctx[(size_t)w]);
        rv_destroy(p->ctx[(size_t)w]);
        p->ctx[(size_t)w] = NULL;
        return rc;
      }
      p->lim[(size_t)w] = L;
      p->ref_seen[(size_t)w] = NULL;
    }
    p->refseq.assign(ref_bases, (size_t)ref_n);
    rv_params P = *params;
    P.candidates_only = P.pileup ? 0 : 1;  // simple-mode text output only prints positions with a passing variant
    std::vector<std::string> ctsv((size_t)n_chunks), cerr((size_t)n_chunks);
    std::vector<BatchTiming> ctm((size_t)n_chunks);
    std::vector<int> crc((size_t)n_chunks, RV_OK);
    std::atomic<int> next(0);
    const int host_thr = std::max(1, host_threads() / p->n_workers);
    const double t_begin = now_ms();
    std::vector<std::thread> th;
    for (int w = 0; w < p->n_workers; ++w)
      th.emplace_back([&, w]() {
        host_threads_override() = host_thr;
        rv_ctx* ctx = p->ctx[(size_t)w];
        rv_set_params(ctx, &P);
        bool first = true;  // the reference slice is uploaded once per worker and run
        for (;;) {
          const int c = next.fetch_add(1);
          if (c >= n_chunks) break;
          const int r0 = cb[(size_t)c], r1 = cb[(size_t)c + 1];
          std::vector<rv_region> regs(regions + r0, regions + r1);
          if (paired) regs.insert(regs.end(), regions + n_regions + r0, regions + n_regions + r1);
          std::vector<std::string> genes((size_t)(r1 - r0), std::string(chr));
          const int flags = 2 | (first ? 1 : 0);
          first = false;
          const int64_t* range = c_rng[(size_t)c].data();
          crc[(size_t)c] = paired ? run_batch_somatic(ctx, *params, B, regs, genes, p->refseq, ref_lo, sample, chr, flags, halo,
                                                      &ctsv[(size_t)c], &ctm[(size_t)c], &cerr[(size_t)c], range)
                                  : run_batch_simple(ctx, P, B, regs, genes, p->refseq, ref_lo, sample, chr, flags, halo,
                                                     &ctsv[(size_t)c], &ctm[(size_t)c], &cerr[(size_t)c], range);
          if (crc[(size_t)c] != RV_OK) break;
        }
        rv_set_params(ctx, params);
      });
    for (size_t i = 0; i < th.size(); ++i) th[i].join();
    const double t_end = now_ms();
    // the chunks' text, in chunk order, into one buffer that keeps its size across runs (no re-zeroing); the copies
    // run on a few threads
    std::vector<size_t> toff((size_t)n_chunks + 1, 0);
    for (int c = 0; c < n_chunks; ++c) {
      if (crc[(size_t)c] != RV_OK) { g_err = cerr[(size_t)c]; return crc[(size_t)c]; }
      toff[(size_t)c + 1] = toff[(size_t)c] + ctsv[(size_t)c].size();
    }
    const size_t total = toff[(size_t)n_chunks];
    if (p->tsv.size() < total + 1) p->tsv.resize(total + 1);
    {
      const int nt = std::max(1, std::min(8, host_threads()));
      std::atomic<int> nc(0);
      std::vector<std::thread> ct;
      char* dst = &p->tsv[0];
      for (int t = 0; t < nt; ++t)
        ct.emplace_back([&]() {
          for (;;) {
            const int c = nc.fetch_add(1);
            if (c >= n_chunks) break;
            memcpy(dst + toff[(size_t)c], ctsv[(size_t)c].data(), ctsv[(size_t)c].size());
          }
        });
      for (size_t i = 0; i < ct.size(); ++i) ct[i].join();
      dst[total] = 0;
    }
    rvh_timing tt;
    memset(&tt, 0, sizeof tt);
    for (int c = 0; c < n_chunks; ++c) {
      const BatchTiming& t = ctm[(size_t)c];
      tt.push_ms += t.push_ms; tt.pileup_ms += t.pileup_ms; tt.fetch_ms += t.fetch_ms; tt.host_ms += t.host_ms;
      tt.patch_ms += t.patch_ms; tt.score_ms += t.score_ms; tt.assemble_ms += t.assemble_ms;
      tt.pileup_kernel_ms += t.pileup_kernel_ms; tt.score_kernel_ms += t.score_kernel_ms;
      tt.n_items += t.stats.n_items; tt.n_reads_kept += t.stats.n_reads_kept; tt.n_aligned_bases += t.stats.n_aligned_bases;
      tt.n_events += t.stats.n_events; tt.n_unsupported += t.stats.n_unsupported; tt.n_variants += t.n_variants;
      tt.n_lines += t.n_lines; tt.h2d_bytes += t.h2d_bytes; tt.d2h_bytes += t.d2h_bytes;
    }
    if (getenv("RV_PIPE_TRACE") && paired) {
      FILE* tf = fopen(getenv("RV_PIPE_TRACE"), "w");
      if (tf) {
        fprintf(tf, "chunk,tiles,t0,t1,t2,t5,t6,t7,fetch_ms,host_ms,patch_ms,h2d_bytes\n");
        for (int c = 0; c < n_chunks; ++c) {
          const BatchTiming& t = ctm[(size_t)c];
          fprintf(tf, "%d,%d,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%lld\n", c, cb[(size_t)c + 1] - cb[(size_t)c],
                  t.t_abs[0] - t_begin, t.t_abs[1] - t_begin, t.t_abs[2] - t_begin, t.t_abs[3] - t_begin, t.t_abs[4] - t_begin,
                  t.t_abs[5] - t_begin, t.fetch_ms, t.host_ms, t.patch_ms, (long long)t.h2d_bytes);
        }
        fclose(tf);
      }
    }
    if (getenv("RV_PIPE_TRACE"))
      fprintf(stderr, "[rvh_pipeline] %d chunks on %d workers: setup %.1f ms, workers %.1f ms, concat %.1f ms\n", n_chunks,
              p->n_workers, t_begin - t_enter, t_end - t_begin, now_ms() - t_end);
    *tsv_out = p->tsv.data();
    *tsv_len = (int64_t)total;
    if (timing) *timing = tt;
    return RV_OK;
  } catch (const std::exception& e) {
    g_err = e.what();
    return RV_ERR_STATE;
  }
}

int rvh_pipeline_run(rvh_pipeline* p, const rv_params* params, const rvh_batch* batch, const rv_region* regions,
                     int32_t n_regions, int32_t chunk_regions, const char* ref_bases, int32_t ref_lo, int64_t ref_n,
                     const char* sample, const char* chr, const char** tsv_out, int64_t* tsv_len, rvh_timing* timing) {
  return pipeline_run(p, params, batch, regions, n_regions, chunk_regions, ref_bases, ref_lo, ref_n, sample, chr, tsv_out,
                      tsv_len, timing, false);
}

int rvh_pipeline_run_paired(rvh_pipeline* p, const rv_params* params, const rvh_batch* batch, const rv_region* regions,
                            int32_t n_regions, int32_t chunk_tiles, const char* ref_bases, int32_t ref_lo, int64_t ref_n,
                            const char* sample, const char* chr, const char** tsv_out, int64_t* tsv_len, rvh_timing* timing) {
  return pipeline_run(p, params, batch, regions, n_regions, chunk_tiles, ref_bases, ref_lo, ref_n, sample, chr, tsv_out,
                      tsv_len, timing, true);
}

const char* rvh_last_error(void) { return g_err.c_str(); }

int64_t rvh_inflate_block(const uint8_t* in, int64_t in_len, uint8_t* out, int64_t out_cap) {
  if (!in || !out || in_len < 0 || out_cap < 0) return -1;
  static thread_local std::unique_ptr<rvio::FastInflate> fi;
  if (!fi) fi.reset(new rvio::FastInflate());
  return (int64_t)fi->inflate(in, (size_t)in_len, out, (size_t)out_cap);
}
uint32_t rvh_crc32(const uint8_t* buf, int64_t n) { return rvio::fast_crc32(buf, (size_t)(n < 0 ? 0 : n)); }

int rvh_run_files(const rv_params* params, const char* fasta, const char* bam, const char* bam2, const char* sample,
                  int32_t n_regions, const char* const* chr, const int32_t* start, const int32_t* end,
                  const char* const* gene, int32_t decode_threads, int32_t gpus, int32_t first_device,
                  char** tsv_out, int64_t* tsv_len, double* cov_info) {
  if (!params || !fasta || !bam || !tsv_out || !tsv_len || n_regions < 0 || (n_regions && (!chr || !start || !end))) return RV_ERR_ARG;
  try {
    FileRunConfig fc;
    fc.fasta = fasta;
    fc.bam = bam;
    fc.bam2 = bam2 ? bam2 : "";
    fc.sample = sample ? sample : "";
    fc.P = *params;
    fc.P.candidates_only = fc.P.pileup ? 0 : 1;
    fc.decode_threads = decode_threads > 0 ? decode_threads : 1;
    fc.gpus = gpus > 0 ? gpus : 1;
    fc.first_device = first_device > 0 ? first_device : 0;
    std::vector<RegionSpec> specs((size_t)n_regions);
    for (int i = 0; i < n_regions; ++i) {
      specs[(size_t)i].chr = chr[i];
      specs[(size_t)i].start = start[i];
      specs[(size_t)i].end = end[i];
      specs[(size_t)i].gene = gene && gene[i] ? gene[i] : chr[i];
    }
    std::string tsv;
    FileRunStats st;
    std::vector<std::string> errors;
    int rc = run_files(fc, specs, &tsv, &st, &errors);
    g_err.clear();
    for (size_t i = 0; i < errors.size(); ++i) g_err += errors[i] + "\n";
    if (rc == 2 && rv_device_count() <= 0) rc = 3;
    char* out = (char*)malloc(tsv.size() + 1);
    if (!out) return RV_ERR_NOMEM;
    memcpy(out, tsv.data(), tsv.size());
    out[tsv.size()] = 0;
    *tsv_out = out;
    *tsv_len = (int64_t)tsv.size();
    if (cov_info) { cov_info[0] = (double)st.cov_sum[0]; cov_info[1] = (double)st.cov_pos[0]; cov_info[2] = (double)st.cov_sum[1]; cov_info[3] = (double)st.cov_pos[1]; }
    return rc;
  } catch (const std::exception& e) {
    g_err = e.what();
    return RV_ERR_STATE;
  }
}
void rvh_free(void* p) { free(p); }


rvh_batch* rvh_load_bam(const char* bam_path, const char* chr, int32_t start, int32_t end, int32_t* chr_len_out) {
  try {
    rvio::SpanScanner rd;
    rvio::BaiIndex bai;
    if (!rd.open(bam_path)) { g_err = std::string("cannot open BAM ") + bam_path; return NULL; }
    if (!bai.load(std::string(bam_path) + ".bai")) { g_err = std::string("cannot open index of ") + bam_path; return NULL; }
    int tid = rd.header().tid_of(chr);
    if (tid < 0) { g_err = std::string("contig not in BAM header: ") + chr; return NULL; }
    if (chr_len_out) *chr_len_out = rd.header().lens[tid];
    rvh_batch* b = new rvh_batch();
    load_span_fast(rd, bai, tid, start, end, &b->b);
    return b;
  } catch (const std::exception& e) {
    g_err = e.what();
    return NULL;
  }
}

int64_t rvh_batch_append(rvh_batch* a, const rvh_batch* b) {
  int64_t off = (int64_t)a->b.reads.size();
  size_t pool_off = (a->b.pool.size() + 15) & ~(size_t)15;
  a->b.pool.resize(pool_off);
  a->b.pool.insert(a->b.pool.end(), b->b.pool.begin(), b->b.pool.end());
  for (size_t i = 0; i < b->b.reads.size(); ++i) {
    rv_read r = b->b.reads[i];
    r.data_off16 += (uint32_t)(pool_off / 16);
    a->b.reads.push_back(r);
  }
  if (b->b.max_ref_span > a->b.max_ref_span) a->b.max_ref_span = b->b.max_ref_span;
  return off;
}
int64_t rvh_batch_n_reads(const rvh_batch* b) { return (int64_t)b->b.reads.size(); }
const rv_read* rvh_batch_reads(const rvh_batch* b) { return b->b.reads.data(); }
const uint8_t* rvh_batch_pool(const rvh_batch* b) { return b->b.pool.data(); }
int64_t rvh_batch_pool_bytes(const rvh_batch* b) { return (int64_t)b->b.pool.size(); }
int32_t rvh_batch_max_ref_span(const rvh_batch* b) { return b->b.max_ref_span; }
int rvh_batch_pin(rvh_batch* b) {
  if (!b) return RV_ERR_ARG;
  if (b->pinned || b->b.reads.empty()) return RV_OK;
  if (cudaHostRegister((void*)b->b.reads.data(), b->b.reads.size() * sizeof(rv_read), cudaHostRegisterDefault) != cudaSuccess ||
      cudaHostRegister((void*)b->b.pool.data(), b->b.pool.size(), cudaHostRegisterDefault) != cudaSuccess) {
    g_err = "cudaHostRegister failed";
    cudaGetLastError();
    return RV_ERR_CUDA;
  }
  b->pinned = true;
  return RV_OK;
}
void rvh_batch_free(rvh_batch* b) { delete b; }

int rvh_make_regions(const rvh_batch* b, const int32_t* starts, const int32_t* ends, int32_t n, int32_t chr_len,
                     int32_t ref_extension, int64_t read_offset, int64_t n_reads_sample, rv_region* out) {
  if (!b || !starts || !ends || !out || n < 0) return RV_ERR_ARG;
  const std::vector<rv_read>& reads = b->b.reads;
  int64_t lo0 = read_offset, hi0 = n_reads_sample < 0 ? (int64_t)reads.size() : read_offset + n_reads_sample;
  for (int i = 0; i < n; ++i) {
    rv_region r;
    r.start = starts[i];
    r.end = ends[i];
    int lo = r.start - ref_extension;
    if (lo < 1) lo = 1;
    int hi = r.end + ref_extension;
    if (hi > chr_len) hi = chr_len;
    r.ref_lo = lo;
    r.ref_hi = hi - 17;
    r.chr_len = chr_len;
    r.max_read_len_in = 0;
    int64_t want_lo = (int64_t)r.start - 1 - b->b.max_ref_span;
    int64_t a = lo0, z = hi0;
    while (a < z) { int64_t m = (a + z) / 2; if ((int64_t)reads[(size_t)m].pos - 1 < want_lo) a = m + 1; else z = m; }
    r.read_lo = a;
    a = lo0; z = hi0;
    while (a < z) { int64_t m = (a + z) / 2; if (reads[(size_t)m].pos - 1 < r.end) a = m + 1; else z = m; }
    r.read_hi = a < r.read_lo ? r.read_lo : a;
    out[i] = r;
  }
  return RV_OK;
}

int64_t rvh_fetch_ref(const char* fasta_path, const char* chr, int32_t lo, int32_t hi, char* out) {
  try {
    rvio::Fasta fa;
    if (!fa.open(fasta_path)) { g_err = std::string("cannot open FASTA/FAI ") + fasta_path; return -1; }
    std::string s;
    if (!fa.fetch(chr, lo, hi, &s)) { g_err = "fetch failed"; return -1; }
    for (size_t i = 0; i < s.size(); ++i) out[i] = (char)toupper((unsigned char)s[i]);
    return (int64_t)s.size();
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

int rvh_install_patch(rv_ctx* ctx, const rv_params* params, const rvh_batch* batch, const rv_region* regions,
                      int32_t n_regions, const char* ref_bases, int32_t ref_lo, int64_t ref_n) {
  if (!ctx || !params || !batch || !regions || !ref_bases) return RV_ERR_ARG;
  try {
    std::vector<rv_region> regs(regions, regions + n_regions);
    std::string refseq(ref_bases, (size_t)ref_n);
    Handoff ho;
    BatchTiming t;
    memset(&t, 0, sizeof t);
    std::string err;
    int rc = host_handoff(ctx, *params, batch->b, regs, refseq, ref_lo, rv_ctx_halo(ctx), &ho, &t, &err);
    if (rc != RV_OK) g_err = err;
    return rc;
  } catch (const std::exception& e) {
    g_err = e.what();
    return RV_ERR_STATE;
  }
}

int rvh_call_regions(rv_ctx* ctx, const rv_params* params, const rvh_batch* batch, const rv_region* regions,
                     int32_t n_regions, const char* ref_bases, int32_t ref_lo, int64_t ref_n, int push_flags,
                     const char* sample, const char* chr, const char** tsv_out, int64_t* tsv_len, rvh_timing* timing) {
  if (!ctx || !params || !batch || !regions || !ref_bases || !tsv_out || !tsv_len) return RV_ERR_ARG;
  try {
    std::vector<rv_region> regs(regions, regions + n_regions);
    std::vector<std::string> genes((size_t)n_regions, std::string(chr));
    std::string refseq(ref_bases, (size_t)ref_n);
    std::string tsv, err;
    BatchTiming t;
    // simple-mode text output only ever prints positions with a passing variant: let the device drop the rest
    rv_params P = *params;
    P.candidates_only = P.pileup ? 0 : 1;
    rv_set_params(ctx, &P);
    int halo = rv_ctx_halo(ctx);
    int rc = run_batch_simple(ctx, P, batch->b, regs, genes, refseq, ref_lo, sample, chr, push_flags, halo, &tsv, &t, &err);
    rv_set_params(ctx, params);
    if (rc != RV_OK) { g_err = err; return rc; }
    std::lock_guard<std::mutex> lk(g_tsv_mu);
    std::string& slot = g_tsv[ctx];
    slot.swap(tsv);
    *tsv_out = slot.data();
    *tsv_len = (int64_t)slot.size();
    if (timing) {
      timing->push_ms = t.push_ms; timing->pileup_ms = t.pileup_ms; timing->fetch_ms = t.fetch_ms;
      timing->host_ms = t.host_ms; timing->patch_ms = t.patch_ms; timing->score_ms = t.score_ms;
      timing->assemble_ms = t.assemble_ms; timing->pileup_kernel_ms = t.pileup_kernel_ms;
      timing->score_kernel_ms = t.score_kernel_ms; timing->n_items = t.stats.n_items;
      timing->n_reads_kept = t.stats.n_reads_kept; timing->n_aligned_bases = t.stats.n_aligned_bases;
      timing->n_events = t.stats.n_events; timing->n_unsupported = t.stats.n_unsupported;
      timing->n_variants = t.n_variants; timing->n_lines = t.n_lines; timing->h2d_bytes = t.h2d_bytes;
      timing->d2h_bytes = t.d2h_bytes;
    }
    return RV_OK;
  } catch (const std::exception& e) {
    g_err = e.what();
    return RV_ERR_STATE;
  }
}

}  // extern "C"
