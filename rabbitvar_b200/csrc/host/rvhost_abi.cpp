// rvhost_abi.cpp — C ABI over the host side of the path (include/rabbitvar_b200_host.h).
#include "../../../include/rabbitvar_b200_host.h"
#include "pipeline.hpp"
#include <map>
#include <mutex>

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

extern "C" {

const char* rvh_last_error(void) { return g_err.c_str(); }

rvh_batch* rvh_load_bam(const char* bam_path, const char* chr, int32_t start, int32_t end, int32_t* chr_len_out) {
  try {
    rvio::BamReader rd;
    rvio::BaiIndex bai;
    if (!rd.open(bam_path)) { g_err = std::string("cannot open BAM ") + bam_path; return NULL; }
    if (!bai.load(std::string(bam_path) + ".bai")) { g_err = std::string("cannot open index of ") + bam_path; return NULL; }
    int tid = rd.header().tid_of(chr);
    if (tid < 0) { g_err = std::string("contig not in BAM header: ") + chr; return NULL; }
    if (chr_len_out) *chr_len_out = rd.header().lens[tid];
    rvh_batch* b = new rvh_batch();
    load_span(rd, bai, tid, start, end, &b->b);
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
