// rv_dump — TEST TOOL.  Runs the pileup(+score) path over a BAM/FASTA for a list of regions and writes the
// same text dump oracle/ref_dump.cpp writes for the reference, so the two can be diffed line by line.
//
//   rv_dump --backend gpu|sim --fasta ref.fa --bam S.bam --chr chrS1 (--region S-E | --bed tiles.bed)
//           [--k 0|1] [--f 0.01] [--u 1] [--three 1] [--fisher 1] [--bam2 1] --out dump.txt [--stages CV]
//
// backend gpu : calls the product through its C ABI (librvgpu.so) — this is what `-m gpu` tests use.
// backend sim : single-steps the SAME __host__ __device__ per-read / per-position functions on the CPU
//               with a plain-add sink.  It exists only so kernel logic can be debugged in a container
//               without a GPU; it is not part of the library and nothing in the product links it.
#include "../../include/rabbitvar_b200.h"
#include "../../rabbitvar_b200/csrc/kernels/rv_core.cuh"
#include "../../rabbitvar_b200/csrc/kernels/rv_score.cuh"
#include "../../rabbitvar_b200/csrc/host/batch_loader.hpp"
#include "../../rabbitvar_b200/csrc/host/pileup_model.hpp"
#include "../../rabbitvar_b200/csrc/host/realign.hpp"
#include "../../rabbitvar_b200/csrc/host/assemble.hpp"
#include <map>
#include <set>
#include <string>
#include <algorithm>

using namespace rvhost;

struct SimSink {
  RegionPileup* R;
  std::vector<rv_event>* events;
  double goodq;
  int64_t kept_reads, kept_bases, unsup, over;
  bool idx(int pos) {
    if (!R->in_table(pos)) { over++; return false; }
    return true;
  }
  long n_direct = 0, n_adj = 0, n_cov = 0, n_segbases = 0;
  bool in_segment = false;
  void single(int pos, int allele, bool dir, int tp, int q, int mapq, int nm) {
    if (!in_segment) n_direct++;
    note_contributor(pos, allele);
    if (getenv("RV_DEBUG_POS") && atoi(getenv("RV_DEBUG_POS")) == pos) fprintf(stderr, "single pos %d al %d dir %d tp %d q %d mapq %d nm %d\n", pos, allele, dir, tp, q, mapq, nm);
    if (!idx(pos)) return;
    uint32_t* row = R->row(pos, allele);
    row[dir ? RV_F_REV : RV_F_FWD] += 1;
    row[RV_F_SUM_TP] += (uint32_t)tp;
    row[RV_F_SUM_Q] += (uint32_t)q;
    row[RV_F_SUM_MAPQ] += (uint32_t)mapq;
    row[RV_F_SUM_NM] += (uint32_t)nm;
    if ((double)q >= goodq) row[RV_F_HI] += 1;
    uint32_t mine = ((uint32_t)tp & 0xffffu) | (((uint32_t)q & 0xffu) << 16) | (1u << 31);
    uint32_t old = row[RV_F_STD];
    if (old == 0) row[RV_F_STD] = mine;
    else {
      if ((old & 0xffffu) != ((uint32_t)tp & 0xffffu)) row[RV_F_STD] |= 1u << 24;
      if (((old >> 16) & 0xffu) != ((uint32_t)q & 0xffu)) row[RV_F_STD] |= 1u << 25;
    }
  }
  // -T: the anchor subtraction is conditional on the row's first contributor (Sink concept, rv_core.cuh); the device defers
  // it to a second pass over the observation list, and so does this harness
  int trim = 0;
  uint32_t cur_ridx = 0;
  std::map<std::pair<int, int>, uint32_t> first_read;
  struct Deferred { int pos, allele; bool dir; int tp, q, mapq, nm; uint32_t ridx; };
  std::vector<Deferred> deferred;
  void note_contributor(int pos, int allele) {
    if (!trim) return;
    std::map<std::pair<int, int>, uint32_t>::iterator it = first_read.find(std::make_pair(pos, allele));
    if (it == first_read.end()) first_read[std::make_pair(pos, allele)] = cur_ridx;
    else if (cur_ridx < it->second) it->second = cur_ridx;
  }
  void sub_anchor(int pos, int allele, bool dir, int tp, int q, int mapq, int nm) {
    if (!trim) { adj(pos, allele, -1, dir, tp, q, mapq, nm); return; }
    Deferred d = {pos, allele, dir, tp, q, mapq, nm, cur_ridx};
    deferred.push_back(d);
  }
  void apply_deferred() {
    for (size_t k = 0; k < deferred.size(); ++k) {
      const Deferred& d = deferred[k];
      std::map<std::pair<int, int>, uint32_t>::const_iterator it = first_read.find(std::make_pair(d.pos, d.allele));
      if (it != first_read.end() && it->second <= d.ridx) adj(d.pos, d.allele, -1, d.dir, d.tp, d.q, d.mapq, d.nm);
    }
    deferred.clear();
    first_read.clear();
  }
  void adj(int pos, int allele, int sign, bool dir, int tp, int q, int mapq, int nm) {
    n_adj++;
    if (sign > 0) note_contributor(pos, allele);
    if (getenv("RV_DEBUG_POS") && atoi(getenv("RV_DEBUG_POS")) == pos) fprintf(stderr, "adj pos %d al %d sign %d dir %d tp %d q %d\n", pos, allele, sign, dir, tp, q);
    if (!idx(pos)) return;
    uint32_t* row = R->row(pos, allele);
    row[dir ? RV_F_REV : RV_F_FWD] += (uint32_t)sign;
    row[RV_F_SUM_TP] += (uint32_t)(sign * tp);
    row[RV_F_SUM_Q] += (uint32_t)(sign * q);
    row[RV_F_SUM_MAPQ] += (uint32_t)(sign * mapq);
    row[RV_F_SUM_NM] += (uint32_t)(sign * nm);
    if ((double)q >= goodq) row[RV_F_HI] += (uint32_t)sign;
  }
  void cov(int pos) {
    if (!in_segment) n_cov++;
    if (!idx(pos)) return;
    R->cov[pos - R->first_pos]++;
  }
  void event(const rv_event& e) { events->push_back(e); }
  // a plain matched stretch (rvk::scan_plain_segment): expanded here the way the gather kernel accumulates it — one
  // single-base observation per base that lies inside the region, the allele being the read base itself
  const rv_read* cur_read;
  const uint8_t* cur_pool;
  int r_start, r_end;
  bool use_segments;
  int scan_segment(const rv_params& P, const rvk::ReadView& rd, const rvk::RefView& ref, int m_start, int rp, int len,
                   bool indel_follows, rvk::SegDesc* out) {
    if (!use_segments) return 0;
    const int kind = rvk::scan_plain_segment(P, rd, ref, m_start, rp, len, indel_follows, out);
    if (ref4) {  // cross-check: the nibble-SIMD proof the kernels use must decide (and list) the same
      int simd = 0;
      rvk::PlainScan ps;
      ps.p_first = ps.p_last = -1;
      const int E0 = m_start - rp - ref.base_pos;
      const int64_t w_lo = ref.lo > ref.base_pos ? ref.lo : ref.base_pos;
      const int64_t w_hi = (int64_t)ref.hi < ref.base_pos + ref.n - 1 ? (int64_t)ref.hi : ref.base_pos + ref.n - 1;
      if (len > 0 && len <= 8192 && E0 >= 0 && m_start >= w_lo && (int64_t)m_start + len - 1 <= w_hi && ((uintptr_t)rd.seq4 & 3) == 0)
        simd = rvk::simd_scan_kind(P, (const uint32_t*)rd.seq4, ref4, E0, rp, len, indel_follows, &ps);
      n_scan++;
      bool same = simd == kind;
      if (same && kind != 0)
        same = (ps.mm_blocks != 0) == (out->mm_blocks != 0) && (uint32_t)ps.ml_lo == out->ml[0] && (uint32_t)(ps.ml_lo >> 32) == out->ml[1] &&
               (uint32_t)ps.ml_hi == out->ml[2] && (uint32_t)(ps.ml_hi >> 32) == out->ml[3] && ps.p_first == out->p_first && ps.p_last == out->p_last;
      if (!same) {
        n_scan_diff++;
        if (getenv("RV_SCAN_DEBUG"))
          fprintf(stderr, "scan diff: pos %d rp %d len %d indel_follows %d scalar %d simd %d n_mm %d/%d flagged [%d,%d]/[%d,%d]\n", m_start, rp, len,
                  (int)indel_follows, kind, simd, out->n_mm, ps.ml_n, out->p_first, out->p_last, ps.p_first, ps.p_last);
      }
      if (kind == 2) n_split++;
    }
    return kind;
  }
  long n_split = 0;
  const uint32_t* ref4 = NULL;
  long n_scan = 0, n_scan_diff = 0;
  bool segment(const rvk::SegDesc& d, bool dir, int mapq, int nm) {
    if (!use_segments) return false;
    const uint8_t* var = cur_pool + (size_t)cur_read->data_off16 * 16;
    const uint8_t* seq4 = var + 4 * (size_t)cur_read->n_cigar;
    const uint8_t* qual = seq4 + ((cur_read->l_seq + 1) >> 1);
    in_segment = true;
    n_segbases += d.len;
    if (getenv("RV_DEBUG_POS") && atoi(getenv("RV_DEBUG_POS")) >= d.m_start && atoi(getenv("RV_DEBUG_POS")) < d.m_start + d.len)
      fprintf(stderr, "segment start %d len %d rp %d re %d rlen %d dir %d nm %d n_mm %d read pos %d\n", d.m_start, d.len, d.rp, d.re, d.rlen, (int)dir, nm, d.n_mm, cur_read->pos);
    for (int k = 0; k < d.len; ++k) {
      const int p = d.m_start + k;
      if (p < r_start || p > r_end) continue;
      const int r = d.rp + k;
      const int b = seq4[r >> 1];
      const int nib = (r & 1) ? (b & 15) : (b >> 4);
      const int al = nib == 1 ? 0 : nib == 2 ? 1 : nib == 4 ? 2 : nib == 8 ? 3 : -1;
      const int up = d.re + k + 1, dn = d.rlen - d.re - k;
      single(p, al, dir, up < dn ? up : dn, qual[r], mapq, nm);
      cov(p);
    }
    in_segment = false;
    return true;
  }
  void max_read_len(int t) { if (t > R->max_read_len) R->max_read_len = t; }
  void kept(int aligned) { kept_reads++; kept_bases += aligned; }
  void unsupported() { unsup++; }
};

struct SimEmit {
  std::vector<rv_variant>* out;
  void emit(const rv_variant& v) { out->push_back(v); }
};

static FILE* OUT = NULL;

static void print_var_fields(const Variation& v) {
  fprintf(OUT, "%d\t%d\t%d\t%.17g\t%.17g\t%.17g\t%.17g\t%d\t%d\t%d\t%d\t%d", v.cnt, v.fwd, v.rev, v.sum_tp, v.sum_q,
          v.sum_mapq, v.sum_nm, v.lo, v.hi, v.pstd ? 1 : 0, v.qstd ? 1 : 0, v.extracnt);
}

static void dump_tables(const char* stage, const RegionPileup& R) {
  // NI: merge dense + sparse, ordered by (pos, key)
  std::map<std::pair<int, std::string>, Variation> all;
  static const char B[5] = "ACGT";
  for (int i = 0; i < R.n_pos; ++i)
    for (int a = 0; a < 4; ++a) {
      const uint32_t* r = R.counts.data() + ((size_t)i * 4 + a) * RV_ROW_U32;
      if (RegionPileup::row_exists(r)) all[std::make_pair(R.first_pos + i, std::string(1, B[a]))] = RegionPileup::row_to_variation(r);
    }
  for (auto& p : R.ni)
    for (auto& k : p.second) all[std::make_pair(p.first, k.first)] = k.second;  // sparse overrides dense
  for (auto& e : all) {
    fprintf(OUT, "%s.NI\t%d\t%s\t", stage, e.first.first, e.first.second.c_str());
    print_var_fields(e.second);
    fputc('\n', OUT);
  }
  for (auto& p : R.ins)
    for (auto& k : p.second) {
      fprintf(OUT, "%s.IN\t%d\t%s\t", stage, p.first, k.first.c_str());
      print_var_fields(k.second);
      fputc('\n', OUT);
    }
  for (int i = 0; i < R.n_pos; ++i)
    if (R.cov[i]) fprintf(OUT, "%s.COV\t%d\t%d\n", stage, R.first_pos + i, (int)R.cov[i]);
  for (int end = 5; end >= 3; end -= 2) {
    const std::map<int, Sclip>& sc = end == 5 ? R.sc5 : R.sc3;
    for (auto& p : sc) {
      fprintf(OUT, "%s.SC%d\t%d\t", stage, end, p.first);
      print_var_fields(p.second);
      fprintf(OUT, "\t%d\n", p.second.used ? 1 : 0);
      for (auto& ie : p.second.nt())
        for (auto& bc : ie.second) fprintf(OUT, "%s.SCNT\t%d\t%d\t%d\t%c\t%d\n", stage, end, p.first, ie.first, bc.first, bc.second);
      for (auto& ie : p.second.seq())
        for (auto& bc : ie.second) {
          fprintf(OUT, "%s.SCSEQ\t%d\t%d\t%d\t%c\t", stage, end, p.first, ie.first, bc.first);
          print_var_fields(bc.second);
          fputc('\n', OUT);
        }
    }
  }
}

static void dump_counts(const char* tag, const std::map<int, std::map<std::string, int> >& m) {
  for (auto& p : m)
    for (auto& k : p.second) fprintf(OUT, "%s\t%d\t%s\t%d\n", tag, p.first, k.first.c_str(), k.second);
}

int main(int argc, char** argv) {
  std::map<std::string, std::string> kv;
  for (int i = 1; i + 1 < argc; i += 2) kv[argv[i]] = argv[i + 1];
  std::string backend = kv.count("--backend") ? kv["--backend"] : "gpu";
  std::string stages = kv.count("--stages") ? kv["--stages"] : "C";
  std::string sample = kv.count("--sample") ? kv["--sample"] : "S";
  rv_params P;
  rv_default_params(&P);
  if (kv.count("--k")) P.local_realign = (uint8_t)atoi(kv["--k"].c_str());
  if (kv.count("--f")) P.freq = atof(kv["--f"].c_str());
  if (kv.count("--u")) P.uniq_u = (uint8_t)atoi(kv["--u"].c_str());
  if (kv.count("--three")) P.move3 = (uint8_t)atoi(kv["--three"].c_str());
  if (kv.count("--fisher")) P.fisher = (uint8_t)atoi(kv["--fisher"].c_str());
  if (kv.count("--UN")) P.uniq_un = (uint8_t)atoi(kv["--UN"].c_str());
  if (kv.count("--Q")) P.mapping_quality = atoi(kv["--Q"].c_str());
  if (kv.count("--m")) P.mismatch = atoi(kv["--m"].c_str());
  if (kv.count("--M")) P.minmatch = atoi(kv["--M"].c_str());
  if (kv.count("--T")) P.trim_bases_after = atoi(kv["--T"].c_str());
  if (kv.count("--q")) P.goodq = atof(kv["--q"].c_str());
  if (kv.count("--X")) P.vext = atoi(kv["--X"].c_str());
  if (kv.count("--B")) P.min_bias_reads = atoi(kv["--B"].c_str());
  if (kv.count("--I")) P.indelsize = atoi(kv["--I"].c_str());
  if (kv.count("--V")) P.lofreq = atof(kv["--V"].c_str());
  if (kv.count("--r")) P.minr = atoi(kv["--r"].c_str());
  if (kv.count("--t")) P.dedup = (uint8_t)atoi(kv["--t"].c_str());
  if (kv.count("--F")) P.samfilter = (int)strtol(kv["--F"].c_str(), NULL, 16);
  if (kv.count("--bam2")) P.has_bam2 = (uint8_t)atoi(kv["--bam2"].c_str());
  if (kv.count("--p")) { P.pileup = 1; P.freq = -1; P.minr = 0; }
  OUT = fopen(kv["--out"].c_str(), "wb");
  if (!OUT) { fprintf(stderr, "rv_dump: cannot open --out\n"); return 2; }

  rvio::BamReader bam;
  rvio::BaiIndex bai;
  rvio::Fasta fa;
  if (!bam.open(kv["--bam"]) || !bai.load(kv["--bam"] + ".bai") || !fa.open(kv["--fasta"])) {
    fprintf(stderr, "rv_dump: cannot open inputs\n");
    return 2;
  }
  std::string chr = kv["--chr"];
  int tid = bam.header().tid_of(chr);
  int32_t chr_len = bam.header().lens[tid];
  std::vector<RegionSpec> specs;
  if (kv.count("--region")) {
    RegionSpec s;
    s.chr = chr;
    sscanf(kv["--region"].c_str(), "%d-%d", &s.start, &s.end);
    specs.push_back(s);
  } else {
    FILE* bf = fopen(kv["--bed"].c_str(), "r");
    char c[256], g[256];
    int s, e;
    while (bf && fscanf(bf, "%255s %d %d %255s", c, &s, &e, g) == 4) {
      RegionSpec r;
      r.chr = c; r.start = s; r.end = e; r.gene = g;
      if (r.chr == chr) specs.push_back(r);
    }
    if (bf) fclose(bf);
  }
  int32_t smin = specs[0].start, smax = specs[0].end;
  for (auto& s : specs) { smin = std::min(smin, s.start); smax = std::max(smax, s.end); }
  ReadBatch batch;
  load_span(bam, bai, tid, smin, smax, &batch);
  std::vector<rv_region> regs;
  make_regions(batch, specs, chr_len, 1200, 0, &regs);
  // reference slice covering every window
  int32_t ref_lo = std::max(1, smin - 1300), ref_hi = std::min(chr_len, smax + 1300);
  std::string refseq;
  fa.fetch(chr, ref_lo, ref_hi, &refseq);
  for (auto& ch : refseq) ch = (char)toupper(ch);

  const int halo = 512;
  std::vector<RegionPileup> rp(regs.size());
  std::vector<rv_event> events;
  std::vector<rv_variant> variants;
  rv_pileup_stats st;
  memset(&st, 0, sizeof st);
  rv_ctx* ctx = NULL;

  if (backend == "sim") {
    rvk::RefView ref;
    ref.bases = refseq.data();
    ref.base_pos = ref_lo;
    ref.n = (int64_t)refseq.size();
    std::vector<uint32_t> ref4((size_t)ref.n / 8 + 16);
    for (size_t w = 0; w < ref4.size(); ++w) ref4[w] = rvk::pack_ref8(refseq.data(), ref.n, (int64_t)w);
    long n_scan = 0, n_scan_diff = 0, n_split = 0;
    for (size_t r = 0; r < regs.size(); ++r) {
      RegionPileup& R = rp[r];
      R.region_idx = (int)r;
      R.start = regs[r].start; R.end = regs[r].end;
      R.first_pos = regs[r].start - halo;
      R.n_pos = regs[r].end - regs[r].start + 1 + 2 * halo;
      R.counts.assign((size_t)R.n_pos * RV_POS_U32, 0);
      R.cov.assign((size_t)R.n_pos, 0);
      R.max_read_len = regs[r].max_read_len_in;
      ref.lo = regs[r].ref_lo;
      ref.hi = regs[r].ref_hi;
      SimSink s;
      s.R = &R; s.events = &events; s.goodq = P.goodq;
      s.trim = P.trim_bases_after;
      s.kept_reads = s.kept_bases = s.unsup = s.over = 0;
      const bool use_fast = getenv("RV_NO_GATHER") == NULL;
      s.use_segments = use_fast && getenv("RV_NO_SEGMENTS") == NULL;
      s.ref4 = ref4.data();
      s.cur_pool = batch.pool.data();
      s.r_start = regs[r].start;
      s.r_end = regs[r].end;
      std::vector<rvk::FastDesc> descs;
      for (int64_t i = regs[r].read_lo; i < regs[r].read_hi; ++i) {
        const rv_read& rd = batch.reads[(size_t)i];
        if (!(rd.pos - 1 < regs[r].end && rd.end_pos > regs[r].start - 1)) continue;
        st.n_items++;
        if (P.dedup && rvk::is_duplicate_read(P, regs[r], batch.reads.data(), batch.pool.data(), i)) continue;
        rvk::FastDesc d;
        memset(&d, 0, sizeof d);
        s.cur_read = &rd;
        s.cur_ridx = (uint32_t)i;
        // RV_SIM_NO_FASTDESC=1: every read goes the way rv_walk_kernel takes (plain stretches through scan_segment)
        static const bool no_fastdesc = getenv("RV_SIM_NO_FASTDESC") != NULL;
        rvk::process_read(P, regs[r], (int)r, rd, batch.pool.data(), ref, (uint32_t)i, s, (use_fast && !no_fastdesc) ? &d : (rvk::FastDesc*)0);
        if (d.m_len) descs.push_back(d);
      }
      s.apply_deferred();
      // the gather kernel's work, position by position
      for (size_t k = 0; k < descs.size(); ++k) {
        const rvk::FastDesc& d = descs[k];
        int lo = std::max(d.m_start, (int)regs[r].start), hi = std::min(d.m_start + (int)d.m_len - 1, (int)regs[r].end);
        for (int p = lo; p <= hi; ++p) {
          int al, tp, q;
          if (!rvk::fast_obs(d, p, batch.pool.data(), &al, &tp, &q)) continue;
          s.single(p, al, d.dir != 0, tp, q, d.mapq, d.nm);
          s.cov(p);
        }
      }
      st.n_reads_kept += s.kept_reads; st.n_aligned_bases += s.kept_bases;
      st.n_unsupported += s.unsup; st.n_overflow += s.over;
      n_scan += s.n_scan; n_scan_diff += s.n_scan_diff; n_split += s.n_split;
      if (getenv("RV_SCAN_DEBUG")) fprintf(stderr, "region %zu: per-base singles %ld, adj %ld, cov %ld, segment bases %ld\n", r, s.n_direct, s.n_adj, s.n_cov, s.n_segbases);
    }
    fprintf(stderr, "rv_dump[sim]: %ld stretches scanned (%ld split around a cluster), %ld where the nibble-SIMD proof and the scalar scan disagree\n", n_scan, n_split, n_scan_diff);
    if (n_scan_diff) return 4;
    st.n_events = (int64_t)events.size();
  } else {
    rv_limits L;
    rv_default_limits(&L);
    L.halo = halo;
    L.max_reads = (int64_t)batch.reads.size() + 16;
    L.max_read_bytes = (int64_t)batch.pool.size() + 64;
    int64_t npos = 0;
    for (auto& r : regs) npos += r.end - r.start + 1 + 2 * halo;
    L.max_positions = npos + 16;
    L.max_regions = (int32_t)regs.size() + 1;
    L.max_events = std::max<int64_t>(1 << 16, (int64_t)batch.reads.size() * 4);
    L.max_variants = npos + 1024;
    L.max_patch = 1 << 20;
    L.max_ref_bases = (int64_t)refseq.size() + 16;
    int rc = rv_create(&ctx, 0, &P, &L);
    if (rc != RV_OK) { fprintf(stderr, "rv_create failed (%d): %s\n", rc, ctx ? rv_last_error(ctx) : "no device"); return 3; }
    rv_read_batch bv = batch.view();
#define RVCK(x) do { int rc_ = (x); if (rc_ != RV_OK) { fprintf(stderr, "%s failed (%d): %s\n", #x, rc_, rv_last_error(ctx)); return 3; } } while (0)
    RVCK(rv_set_reference(ctx, ref_lo, (int64_t)refseq.size(), refseq.data()));
    RVCK(rv_push_reads(ctx, &bv));
    RVCK(rv_set_regions(ctx, regs.data(), (int32_t)regs.size()));
    RVCK(rv_pileup(ctx));
    RVCK(rv_get_pileup_stats(ctx, &st));
    const rv_event* ev;
    int64_t nev;
    RVCK(rv_fetch_events(ctx, &ev, &nev));
    events.assign(ev, ev + nev);
    const int32_t* mrl;
    int32_t nmrl;
    RVCK(rv_fetch_max_read_len(ctx, &mrl, &nmrl));
    for (size_t r = 0; r < regs.size(); ++r) {
      RegionPileup& R = rp[r];
      const uint32_t *c, *cv;
      RVCK(rv_fetch_tables(ctx, (int32_t)r, &c, &cv, &R.first_pos, &R.n_pos));
      R.region_idx = (int)r;
      R.start = regs[r].start; R.end = regs[r].end;
      R.counts.assign(c, c + (size_t)R.n_pos * RV_POS_U32);
      R.cov.assign(cv, cv + R.n_pos);
      R.max_read_len = mrl[r];
    }
  }
  if (backend == "sim") {
    std::stable_sort(events.begin(), events.end(), [](const rv_event& a, const rv_event& b) {
      if (a.region != b.region) return a.region < b.region;
      if (a.read_idx != b.read_idx) return a.read_idx < b.read_idx;
      return a.seq_no < b.seq_no;
    });
  }
  reduce_events(events.data(), (int64_t)events.size(), batch, P.goodq, rp);

  // stage R: host realigner on the pileup model, then patch upload + scoring
  rvk::RefView refv;
  refv.bases = refseq.data();
  refv.base_pos = ref_lo;
  refv.n = (int64_t)refseq.size();
  std::vector<std::vector<rv_patch_entry> > patches(regs.size());
  FILE* const final_out = OUT;
  std::vector<char*> rbuf(regs.size(), (char*)NULL);
  std::vector<size_t> rlen(regs.size(), 0);
  std::vector<FILE*> rfile(regs.size(), (FILE*)NULL);
  for (size_t r = 0; r < regs.size(); ++r) {
    rfile[r] = open_memstream(&rbuf[r], &rlen[r]);
    OUT = rfile[r];
    fprintf(OUT, "REGION\t%s\t%s\t%d\t%d\n", sample.c_str(), chr.c_str(), regs[r].start, regs[r].end);
    if (stages.find('C') != std::string::npos) {
      fprintf(OUT, "C.MAXRL\t%d\n", rp[r].max_read_len);
      dump_tables("C", rp[r]);
      dump_counts("C.PINS", rp[r].pins);
      dump_counts("C.PDEL", rp[r].pdel);
      dump_counts("C.MNP", rp[r].mnp);
    }
    refv.lo = regs[r].ref_lo;
    refv.hi = regs[r].ref_hi;
    realign_region(P, rp[r], refv, regs[r].chr_len, NULL, &batch, regs[r].read_lo, regs[r].read_hi);
    if (stages.find('R') != std::string::npos) {
      fprintf(OUT, "R.MAXRL\t%d\n", rp[r].max_read_len);
      dump_tables("R", rp[r]);
    }
    build_patch(rp[r], &patches[r]);
    if (stages.find('V') != std::string::npos && backend == "sim") {
      // CPU single-step of score_position over the region
      rvk::LgTable lgt;
      lgt.t = NULL; lgt.n = 0;
      RegionPileup& R = rp[r];
      std::vector<rv_patch_entry>& pe = patches[r];
      std::map<int, std::pair<int, int> > grp;
      for (size_t i = 0; i < pe.size();) {
        size_t j = i;
        while (j < pe.size() && pe[j].pos == pe[i].pos) ++j;
        grp[pe[i].pos] = std::make_pair((int)i, (int)(j - i));
        i = j;
      }
      SimEmit em;
      em.out = &variants;
      int unsup = 0;
      for (int pos = regs[r].start; pos <= regs[r].end; ++pos) {
        int i = pos - R.first_pos;
        bool has_next = i + 1 < R.n_pos;
        int pf = 0, pn = 0;
        if (grp.count(pos)) { pf = grp[pos].first; pn = grp[pos].second; }
        int pfn = 0, pnn = 0;
        if (has_next && grp.count(pos + 1)) { pfn = grp[pos + 1].first; pnn = grp[pos + 1].second; }
        rvk::score_position(P, regs[r], (int)r, pos, refv, R.counts.data() + (size_t)i * RV_POS_U32, R.cov[i], has_next,
                            R.counts.data() + (size_t)(i + 1) * RV_POS_U32, has_next ? R.cov[i + 1] : 0u, pe.data(), pf, pn,
                            pfn, pnn, lgt, em, &unsup);
      }
    }
  }
  if (stages.find('V') != std::string::npos && backend != "sim") {
    // upload patches of all regions, score on the device
    std::vector<rv_patch_entry> all;
    std::vector<int32_t> creg, cpos, cval;
    for (size_t r = 0; r < regs.size(); ++r) {
      for (auto& e : patches[r]) all.push_back(e);
      collect_cov_patch(rp[r], &creg, &cpos, &cval);
    }
    RVCK(rv_apply_patch(ctx, all.data(), (int64_t)all.size(), creg.data(), cpos.data(), cval.data(), (int64_t)creg.size()));
    RVCK(rv_score(ctx));
    const rv_variant* vv;
    int64_t nv;
    RVCK(rv_fetch_variants(ctx, &vv, &nv));
    variants.assign(vv, vv + nv);
    // key_id of patch entries is an index into the concatenated list: rebase per region below
    size_t base = 0;
    std::vector<size_t> bases;
    for (size_t r = 0; r < regs.size(); ++r) { bases.push_back(base); base += patches[r].size(); }
    for (auto& v : variants)
      if (v.key_kind == 1) v.key_id -= (int32_t)bases[v.region];
  }
  if (stages.find('V') != std::string::npos) {
    std::stable_sort(variants.begin(), variants.end(), [](const rv_variant& a, const rv_variant& b) {
      if (a.region != b.region) return a.region < b.region;
      if (a.pos != b.pos) return a.pos < b.pos;
      return a.rank < b.rank;
    });
    // V lines go to their own region's buffer
    for (size_t i = 0; i < variants.size();) {
      size_t j = i;
      while (j < variants.size() && variants[j].region == variants[i].region) ++j;
      std::vector<rv_variant> one(variants.begin() + i, variants.begin() + j);
      dump_variants(rfile[(size_t)variants[i].region], P, one, patches, regs, refv, chr);
      i = j;
    }
  }
  for (size_t r = 0; r < regs.size(); ++r) {
    fclose(rfile[r]);
    fwrite(rbuf[r], 1, rlen[r], final_out);
    free(rbuf[r]);
  }
  OUT = final_out;
  fprintf(stderr, "rv_dump[%s]: items %lld kept %lld bases %lld events %lld overflow %lld unsupported %lld variants %zu\n",
          backend.c_str(), (long long)st.n_items, (long long)st.n_reads_kept, (long long)st.n_aligned_bases,
          (long long)st.n_events, (long long)st.n_overflow, (long long)st.n_unsupported, variants.size());
  if (ctx) rv_destroy(ctx);
  fclose(OUT);
  return 0;
}
