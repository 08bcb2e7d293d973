// batch_loader.hpp — host decode path: BGZF/BAM -> rv_read headers + packed byte pool (the staging
// buffers rv_push_reads uploads), plus the per-region read ranges.
//
// Replaces RecordPreprocessor's iterator construction (reference src/recordPreprocessor.cpp:10-32) and
// RecordPreprocessor::makeReference (:41-78): the region list of a batch is served from ONE sequential
// scan of the BAM over the union of the regions; each region then refers to a contiguous slice
// [read_lo, read_hi) of that scan, and the device applies the htslib overlap test per (region, read).
#pragma once
#include "../../../include/rabbitvar_b200.h"
#include "../io/bamio.hpp"
#include <string>
#include <vector>

namespace rvhost {

struct RegionSpec {
  std::string chr;
  int32_t start, end;  // 1-based inclusive
  std::string gene;
};

struct ReadBatch {
  std::vector<rv_read> reads;
  std::vector<uint8_t> pool;
  int32_t max_ref_span;
  ReadBatch() : max_ref_span(0) {}
  void clear() { reads.clear(); pool.clear(); max_ref_span = 0; }
  void swap(ReadBatch& o) { reads.swap(o.reads); pool.swap(o.pool); std::swap(max_ref_span, o.max_ref_span); }
  rv_read_batch view() const {
    rv_read_batch b;
    b.n_reads = (int64_t)reads.size();
    b.reads = reads.data();
    b.pool = pool.data();
    b.pool_bytes = (int64_t)pool.size();
    return b;
  }
  const uint32_t* cigar(size_t i) const { return (const uint32_t*)(pool.data() + (size_t)reads[i].data_off16 * 16); }
  const uint8_t* seq4(size_t i) const { return pool.data() + (size_t)reads[i].data_off16 * 16 + 4 * (size_t)reads[i].n_cigar; }
  const uint8_t* qual(size_t i) const { return seq4(i) + ((reads[i].l_seq + 1) >> 1); }
  char base(size_t i, int k) const {
    int b = seq4(i)[k >> 1];
    return "=ACMGRSVTWYHKDBN"[(k & 1) ? (b & 15) : (b >> 4)];
  }
};

inline void append_record(ReadBatch& out, const rvio::BamRecord& r) {
  rv_read h;
  h.pos = r.pos + 1;
  h.mpos = r.mpos + 1;
  h.l_seq = r.l_seq;
  h.flag = r.flag;
  h.n_cigar = r.n_cigar;
  int64_t nm;
  h.nm = rvio::aux_get_int(r.aux(), r.aux_len(), "NM", &nm) ? (int16_t)nm : (int16_t)-1;
  h.mapq = r.mapq;
  h.mate_same_tid = r.tid == r.mtid ? 1 : 0;
  h.end_pos = r.end_pos();
  h.mtid = r.mtid;
  size_t bytes = 4 * (size_t)r.n_cigar + (size_t)((r.l_seq + 1) >> 1) + (size_t)r.l_seq;
  size_t off = (out.pool.size() + 15) & ~(size_t)15;
  if ((off >> 4) > (size_t)0xffffffffu) throw std::runtime_error("read batch pool exceeds 64 GiB: use smaller region blocks");
  h.data_off16 = (uint32_t)(off / 16);
  out.pool.resize(off + bytes);
  if (bytes) memcpy(out.pool.data() + off, r.data.data() + r.l_qname, bytes);
  int span = h.end_pos - r.pos;
  if (span > out.max_ref_span) out.max_ref_span = span;
  out.reads.push_back(h);
}

// The same from a record that still lies in the inflated BGZF stream (rvio::SpanScanner): one memcpy of the
// record's cigar | seq | qual bytes into the pool, no intermediate copy.
inline void append_raw_record(ReadBatch& out, const rvio::RawRecord& r) {
  rv_read h;
  h.pos = r.pos() + 1;
  h.mpos = r.mpos() + 1;
  h.l_seq = r.l_seq();
  h.flag = r.flag();
  h.n_cigar = r.n_cigar();
  int64_t nm;
  h.nm = rvio::aux_get_int(r.aux(), r.aux_len(), "NM", &nm) ? (int16_t)nm : (int16_t)-1;
  h.mapq = r.mapq();
  h.mate_same_tid = r.tid() == r.mtid() ? 1 : 0;
  h.end_pos = r.end_pos();
  h.mtid = r.mtid();
  const size_t bytes = r.tail_bytes();
  const size_t off = (out.pool.size() + 15) & ~(size_t)15;
  if ((off >> 4) > (size_t)0xffffffffu) throw std::runtime_error("read batch pool exceeds 64 GiB: use smaller region blocks");
  h.data_off16 = (uint32_t)(off / 16);
  if (out.pool.size() != off) out.pool.resize(off);  // (padding of the previous record)
  out.pool.insert(out.pool.end(), r.cigar_bytes(), r.cigar_bytes() + bytes);
  const int span = h.end_pos - r.pos();
  if (span > out.max_ref_span) out.max_ref_span = span;
  out.reads.push_back(h);
}

// Appends every read overlapping [span_start, span_end] of `tid`, in file order, to `out` (which keeps what it
// holds: the spans of one job are loaded one after the other, ascending).
inline void load_span_fast(rvio::SpanScanner& sc, const rvio::BaiIndex& bai, int tid, int32_t span_start, int32_t span_end,
                           ReadBatch* out) {
  sc.scan(bai, tid, (int64_t)span_start - 1, (int64_t)span_end, [&](const rvio::RawRecord& r) { append_raw_record(*out, r); });
  out->pool.resize((out->pool.size() + 15) & ~(size_t)15);
}

// Loads every read overlapping [span_start, span_end] of `tid` in file order.
inline bool load_span(rvio::BamReader& rd, const rvio::BaiIndex& bai, int tid, int32_t span_start, int32_t span_end,
                      ReadBatch* out) {
  out->clear();
  rvio::BamRegionIter it;
  it.start(&rd, &bai, tid, (int64_t)span_start - 1, (int64_t)span_end);
  rvio::BamRecord rec;
  while (it.next(rec)) append_record(*out, rec);
  out->pool.resize((out->pool.size() + 15) & ~(size_t)15);
  return true;
}

// Concatenates b onto a (the reads of a second sample); returns the read-index offset of b's reads inside a.
inline int64_t append_batch(ReadBatch& a, const ReadBatch& b) {
  const int64_t off = (int64_t)a.reads.size();
  const size_t pool_off = (a.pool.size() + 15) & ~(size_t)15;
  a.pool.resize(pool_off);
  a.pool.insert(a.pool.end(), b.pool.begin(), b.pool.end());
  for (size_t i = 0; i < b.reads.size(); ++i) {
    rv_read r = b.reads[i];
    r.data_off16 += (uint32_t)(pool_off / 16);
    a.reads.push_back(r);
  }
  if (b.max_ref_span > a.max_ref_span) a.max_ref_span = b.max_ref_span;
  return off;
}

// Region descriptors against reads [r0, r1) of the batch only (one fetched span of one sample; the batch may hold
// several spans / samples one after the other, each sorted by position).
inline void make_regions_range(const ReadBatch& b, const std::vector<RegionSpec>& specs, int32_t chr_len, int32_t ref_ext,
                               int32_t nucl_ext, int64_t r0, int64_t r1, std::vector<rv_region>* out) {
  out->clear();
  for (size_t i = 0; i < specs.size(); ++i) {
    rv_region r;
    r.start = specs[i].start;
    r.end = specs[i].end;
    int lo = r.start - nucl_ext - ref_ext;
    if (lo < 1) lo = 1;
    int hi = r.end + nucl_ext + ref_ext;
    if (hi > chr_len) hi = chr_len;
    r.ref_lo = lo;
    r.ref_hi = hi - 17;
    r.chr_len = chr_len;
    r.max_read_len_in = 0;
    const int64_t want_lo = (int64_t)r.start - 1 - b.max_ref_span;
    int64_t a = r0, z = r1;
    while (a < z) { const int64_t m = (a + z) / 2; if ((int64_t)b.reads[(size_t)m].pos - 1 < want_lo) a = m + 1; else z = m; }
    r.read_lo = a;
    a = r0; z = r1;
    while (a < z) { const int64_t m = (a + z) / 2; if (b.reads[(size_t)m].pos - 1 < r.end) a = m + 1; else z = m; }
    r.read_hi = a < r.read_lo ? r.read_lo : a;
    out->push_back(r);
  }
}

// Fills rv_region entries for regions (all on one contig, any order) against a loaded batch.
// ref window = [max(1, start - x - Y), min(len, end + x + Y) - 17]  (recordPreprocessor.cpp:42-55,
// CONF_SEED_1 = 17); here x (numberNucleotideToExtend) is already applied to the region by the caller.
inline void make_regions(const ReadBatch& b, const std::vector<RegionSpec>& specs, int32_t chr_len, int32_t ref_ext,
                         int32_t nucl_ext, std::vector<rv_region>* out) {
  out->clear();
  const size_t n = b.reads.size();
  for (size_t i = 0; i < specs.size(); ++i) {
    rv_region r;
    r.start = specs[i].start;
    r.end = specs[i].end;
    int lo = r.start - nucl_ext - ref_ext;
    if (lo < 1) lo = 1;
    int hi = r.end + nucl_ext + ref_ext;
    if (hi > chr_len) hi = chr_len;
    r.ref_lo = lo;
    r.ref_hi = hi - 17;
    r.chr_len = chr_len;
    r.max_read_len_in = 0;
    // reads are sorted by pos: [first read with pos0 >= beg0 - max_span, first read with pos0 >= end)
    int64_t want_lo = (int64_t)r.start - 1 - b.max_ref_span;
    size_t a = 0, z = n;
    while (a < z) { size_t m = (a + z) / 2; if ((int64_t)b.reads[m].pos - 1 < want_lo) a = m + 1; else z = m; }
    r.read_lo = (int64_t)a;
    a = 0; z = n;
    while (a < z) { size_t m = (a + z) / 2; if (b.reads[m].pos - 1 < r.end) a = m + 1; else z = m; }
    r.read_hi = (int64_t)a;
    if (r.read_hi < r.read_lo) r.read_hi = r.read_lo;
    out->push_back(r);
  }
}

}  // namespace rvhost
