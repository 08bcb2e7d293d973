// pipeline.hpp — the product's replacement for one_region_run (reference src/modes/simpleMode.cpp:18-64)
// over a BATCH of regions: host buffers in, TSV text out.
//
//   rv_push_reads -> rv_pileup (GPU)            == CigarParser::process
//   events/tables D2H -> reduce -> realign      == VariationRealigner::process (host, north_star)
//   rv_apply_patch -> rv_score (GPU)            == ToVarsBuilder::process
//   variants D2H -> assemble/filter/format      == SimpleMode::output + print_output_variant_simple
#pragma once
#include "batch_loader.hpp"
#include "pileup_model.hpp"
#include "realign.hpp"
#include "assemble.hpp"
#include "somatic.hpp"
#include <chrono>

namespace rvhost {

struct BatchTiming {
  double push_ms, pileup_ms, fetch_ms, host_ms, patch_ms, score_ms, assemble_ms;
  float pileup_kernel_ms, score_kernel_ms;
  rv_pileup_stats stats;
  int64_t n_variants, n_lines, h2d_bytes, d2h_bytes;
  int64_t cov_sum[2], cov_pos[2];  // paired mode: coverage summary of the tumor / normal tiles (<out>.info)
  double t_abs[6];                 // paired mode: stage boundaries on the host clock (RV_PIPE_TRACE)
};

inline double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Handoff {
  std::vector<std::vector<rv_patch_entry> > patches;  // per region
  std::vector<size_t> bases;                          // offset of each region's entries in the uploaded list
};

// After rv_pileup: events + dense tables D2H, BAM-order reduce of the sparse keys, host realigner,
// write-back of the result as the patch list (rv_apply_patch).
inline int host_handoff(rv_ctx* ctx, const rv_params& P, const ReadBatch& batch, const std::vector<rv_region>& regs,
                        const std::string& refseq, int32_t ref_lo, int halo, Handoff* out, BatchTiming* t,
                        std::string* err, const std::vector<int>* pair_of = NULL) {
  int rc;
#define RV_STEP(x)                                                  \
  do {                                                              \
    rc = (x);                                                       \
    if (rc != RV_OK) {                                              \
      if (err) *err = std::string(#x) + ": " + rv_last_error(ctx);  \
      return rc;                                                    \
    }                                                               \
  } while (0)
  double t2 = now_ms();
  const rv_event* ev;
  int64_t nev;
  RV_STEP(rv_fetch_events(ctx, &ev, &nev));
  double t2a = now_ms();
  const int32_t* mrl;
  int32_t nmrl;
  RV_STEP(rv_fetch_max_read_len(ctx, &mrl, &nmrl));
  std::vector<RegionPileup> rp(regs.size());
  for (size_t r = 0; r < regs.size(); ++r) {
    RegionPileup& R = rp[r];
    R.region_idx = (int)r;
    R.start = regs[r].start;
    R.end = regs[r].end;
    R.first_pos = regs[r].start - halo;
    R.n_pos = regs[r].end - regs[r].start + 1 + 2 * halo;
    R.dense = false;
    R.max_read_len = mrl[r];
    // somatic mode: the normal pass starts from the tumor's maxReadLength (somaticMode.cpp:109)
    if (pair_of && (*pair_of)[r] >= 0 && mrl[(*pair_of)[r]] > R.max_read_len) R.max_read_len = mrl[(*pair_of)[r]];
  }
  // dense rows the host stage will look at: around multi-nucleotide keys (adjustMNP reads the single-base
  // keys left/right of an MNV, VariationRealigner.cpp:357-399) and under TTREF observations
  {
    std::vector<int32_t> qreg, qpos;
    std::set<std::pair<int, int> > seen;
    for (int64_t i = 0; i < nev; ++i) {
      const rv_event& e = ev[i];
      int lo = 0, hi = -1;
      if (e.kind == RV_EV_NI && (e.flags & RV_EVF_MNP)) { lo = e.pos - 1; hi = e.pos + e.keylen + 1; }
      else if (e.kind == RV_EV_TTREF) { lo = hi = e.pos; }
      else if (P.local_realign && e.kind == RV_EV_NI && (e.flags & RV_EVF_PDEL)) {
        // realigndel looks at the reference allele under the deletion and at mismatches right next to it
        // (a first guess: whatever else it touches is fetched by the re-run loop below)
        int dellen = 0;
        for (int k = 1; k < e.keylen && e.key[k] >= '0' && e.key[k] <= '9'; ++k) dellen = dellen * 10 + (e.key[k] - '0');
        lo = e.pos - 3;
        hi = e.pos + dellen + 4;
      } else if (P.local_realign && e.kind == RV_EV_IN) { lo = e.pos - 3; hi = e.pos + 5; }
      for (int p = lo; p <= hi; ++p)
        if (seen.insert(std::make_pair(e.region, p)).second) { qreg.push_back(e.region); qpos.push_back(p); }
    }
    const uint32_t* rows = NULL;
    double t2b = now_ms();
    RV_STEP(rv_fetch_rows(ctx, qreg.data(), qpos.data(), (int64_t)qreg.size(), &rows));
    if (getenv("RV_TRACE"))
      fprintf(stderr, "[trace] fetch_events %.1f ms (%lld), interest set %.1f ms (%zu rows), fetch_rows %.1f ms\n", t2a - t2,
              (long long)nev, t2b - t2a, qreg.size(), now_ms() - t2b);
    for (size_t i = 0; i < qreg.size(); ++i) {
      std::array<uint32_t, 33> a;
      memcpy(a.data(), rows + i * 33, sizeof(uint32_t) * 33);
      rp[(size_t)qreg[i]].srows[qpos[i]] = a;
    }
    t->d2h_bytes += (int64_t)qreg.size() * 33 * 4;
  }
  t->d2h_bytes += nev * (int64_t)sizeof(rv_event);
  double t3 = now_ms();
  reduce_events(ev, nev, batch, P.goodq, rp);
  rvk::RefView refv;
  refv.bases = refseq.data();
  refv.base_pos = ref_lo;
  refv.n = (int64_t)refseq.size();
  out->patches.assign(regs.size(), std::vector<rv_patch_entry>());
  out->bases.clear();
  std::vector<rv_patch_entry> all;
  std::vector<int32_t> creg, cpos, cval;
  std::vector<int> bad(regs.size(), 0);
  // The realigner reads dense rows next to indels and soft clips that cannot be known in advance.  It runs on a
  // copy of the region state; rows it touched without having them are fetched and the region is re-run from the
  // pristine state (the stage is a pure function of its inputs) until a run completes without a miss.
  {
    std::vector<RegionPileup> work(regs.size());
    std::vector<size_t> todo(regs.size());
    for (size_t r = 0; r < regs.size(); ++r) todo[r] = r;
    for (int iter = 0; !todo.empty(); ++iter) {
      parallel_for(todo.size(), host_threads(), [&](size_t k) {
        const size_t r = todo[k];
        work[r] = rp[r];
        rvk::RefView rv = refv;
        rv.lo = regs[r].ref_lo;
        rv.hi = regs[r].ref_hi;
        realign_region(P, work[r], rv, regs[r].chr_len, NULL, &batch, regs[r].read_lo, regs[r].read_hi);
      });
      std::vector<int32_t> qreg, qpos;
      std::vector<size_t> again;
      for (size_t k = 0; k < todo.size(); ++k) {
        const size_t r = todo[k];
        if (!work[r].row_misses) continue;
        again.push_back(r);
        for (std::map<int, std::array<uint32_t, 33> >::const_iterator it = work[r].srows.begin(); it != work[r].srows.end(); ++it)
          if (!rp[r].srows.count(it->first)) { qreg.push_back((int32_t)r); qpos.push_back(it->first); }
      }
      if (again.empty()) break;
      if (iter >= 8) {
        if (err) *err = "host stage did not converge on the set of dense rows it needs";
        return RV_ERR_STATE;
      }
      const uint32_t* rows = NULL;
      RV_STEP(rv_fetch_rows(ctx, qreg.data(), qpos.data(), (int64_t)qreg.size(), &rows));
      for (size_t i = 0; i < qreg.size(); ++i) {
        std::array<uint32_t, 33> a;
        memcpy(a.data(), rows + i * 33, sizeof(uint32_t) * 33);
        rp[(size_t)qreg[i]].srows[qpos[i]] = a;
      }
      t->d2h_bytes += (int64_t)qreg.size() * 33 * 4;
      todo.swap(again);
    }
    parallel_for(regs.size(), host_threads(), [&](size_t r) {
      rp[r] = work[r];
      if (rp[r].row_misses) bad[r] = 1;
      build_patch(rp[r], &out->patches[r]);
    });
  }
  for (size_t r = 0; r < regs.size(); ++r) {
    if (bad[r]) {
      if (err) *err = "host stage touched dense rows that were not fetched";
      return RV_ERR_STATE;
    }
    out->bases.push_back(all.size());
    all.insert(all.end(), out->patches[r].begin(), out->patches[r].end());
    collect_cov_patch(rp[r], &creg, &cpos, &cval);
  }
  double t4 = now_ms();
  RV_STEP(rv_apply_patch(ctx, all.data(), (int64_t)all.size(), creg.data(), cpos.data(), cval.data(), (int64_t)creg.size()));
  t->h2d_bytes += (int64_t)all.size() * (int64_t)sizeof(rv_patch_entry) + (int64_t)creg.size() * 12;
  double t5 = now_ms();
  t->fetch_ms = t3 - t2;
  t->host_ms = t4 - t3;
  t->patch_ms = t5 - t4;
#undef RV_STEP
  return RV_OK;
}

// Runs every region of `regs` (all on one contig) through the path; appends the TSV lines of simple mode.
// `genes[i]` is the BED name column of region i.  Returns an rv_* error code.
inline int run_batch_simple(rv_ctx* ctx, const rv_params& P, const ReadBatch& batch, const std::vector<rv_region>& regs,
                            const std::vector<std::string>& genes, const std::string& refseq, int32_t ref_lo,
                            const std::string& sample, const std::string& chr, int push_flags, int halo,
                            std::string* tsv, BatchTiming* tm, std::string* err, const int64_t* read_range = NULL) {
  BatchTiming t;
  memset(&t, 0, sizeof t);
  int rc;
#define RV_STEP(x)                                                  \
  do {                                                              \
    rc = (x);                                                       \
    if (rc != RV_OK) {                                              \
      if (err) *err = std::string(#x) + ": " + rv_last_error(ctx);  \
      return rc;                                                    \
    }                                                               \
  } while (0)
  double t0 = now_ms();
  const bool push_reference = (push_flags & 1) != 0, push_reads = (push_flags & 2) != 0;
  if (push_reference) RV_STEP(rv_set_reference(ctx, ref_lo, (int64_t)refseq.size(), refseq.data()));
  if (push_reads) {
    rv_read_batch bv = batch.view();
    if (read_range) {
      RV_STEP(rv_push_reads_range(ctx, &bv, read_range[0], read_range[1]));
      if (read_range[1] > read_range[0]) {
        const int64_t p_lo = (int64_t)batch.reads[(size_t)read_range[0]].data_off16 * 16;
        const int64_t p_hi = read_range[1] < (int64_t)batch.reads.size()
                                 ? (int64_t)batch.reads[(size_t)read_range[1]].data_off16 * 16 : (int64_t)batch.pool.size();
        t.h2d_bytes += (read_range[1] - read_range[0]) * (int64_t)sizeof(rv_read) + (p_hi - p_lo);
      }
    } else {
      RV_STEP(rv_push_reads(ctx, &bv));
      t.h2d_bytes += (int64_t)batch.reads.size() * (int64_t)sizeof(rv_read) + (int64_t)batch.pool.size();
    }
  }
  RV_STEP(rv_set_regions(ctx, regs.data(), (int32_t)regs.size()));
  t.h2d_bytes += (push_reference ? (int64_t)refseq.size() : 0);
  double t1 = now_ms();
  RV_STEP(rv_pileup(ctx));
  RV_STEP(rv_get_pileup_stats(ctx, &t.stats));
  double t2 = now_ms();
  Handoff ho;
  rc = host_handoff(ctx, P, batch, regs, refseq, ref_lo, halo, &ho, &t, err);
  if (rc != RV_OK) return rc;
  std::vector<std::vector<rv_patch_entry> >& patches = ho.patches;
  std::vector<size_t>& bases = ho.bases;
  rvk::RefView refv;
  refv.bases = refseq.data();
  refv.base_pos = ref_lo;
  refv.n = (int64_t)refseq.size();
  double t5 = now_ms();
  RV_STEP(rv_score(ctx));
  const rv_variant* vv;
  int64_t nv;
  RV_STEP(rv_fetch_variants(ctx, &vv, &nv));
  t.d2h_bytes += nv * (int64_t)sizeof(rv_variant);
  double t6 = now_ms();
  // records are sorted by (region, pos, rank): format every region's lines independently, concatenate in order
  std::vector<int64_t> rfirst(regs.size() + 1, nv);
  {
    int64_t i = 0;
    for (size_t r = 0; r < regs.size(); ++r) {
      while (i < nv && vv[i].region < (int)r) ++i;
      rfirst[r] = i;
    }
    rfirst[regs.size()] = nv;
  }
  std::vector<std::string> rtsv(regs.size());
  std::vector<int64_t> rlines(regs.size(), 0);
  parallel_for(regs.size(), host_threads(), [&](size_t r) {
    rvk::RefView rv = refv;
    rv.lo = regs[r].ref_lo;
    rv.hi = regs[r].ref_hi;
    std::vector<rv_variant> group;
    std::string& out_s = rtsv[r];
    for (int64_t i = rfirst[r]; i < rfirst[r + 1];) {
      int64_t j = i;
      while (j < rfirst[r + 1] && vv[j].pos == vv[i].pos) ++j;
      if (j - i == 1 && vv[i].is_ref && !P.pileup) { i = j; continue; }  // SimpleMode::output skips variant-less positions
      group.assign(vv + i, vv + j);
      for (size_t k = 0; k < group.size(); ++k)
        if (group[k].key_kind == 1) group[k].key_id -= (int32_t)bases[r];
      PositionVars pv;
      assemble_position(P, group.data(), (int)group.size(), patches[r], rv, regs[r].chr_len, &pv);
      size_t before = out_s.size();
      output_position_simple(P, pv, sample, genes[r], chr, regs[r].start, regs[r].end, &out_s);
      for (size_t k = before; k < out_s.size(); ++k) rlines[r] += out_s[k] == '\n';
      i = j;
    }
  });
  for (size_t r = 0; r < regs.size(); ++r) {
    tsv->append(rtsv[r]);
    t.n_lines += rlines[r];
  }
  double t7 = now_ms();
  t.push_ms = t1 - t0; t.pileup_ms = t2 - t1;
  t.score_ms = t6 - t5; t.assemble_ms = t7 - t6;
  t.n_variants = nv;
  rv_last_kernel_ms(ctx, &t.pileup_kernel_ms, &t.score_kernel_ms);
  if (tm) *tm = t;
#undef RV_STEP
  return RV_OK;
}

// Paired (tumor | normal) form of run_batch_simple: regs = the n tumor tiles followed by the same n tiles of the
// normal sample (one_region_run_somt + SomaticMode::output, somaticMode.cpp:83-127, :311-352).  Both samples
// are piled in one launch; scoring runs twice: once per sample with the device-side candidate cut to find the
// positions where either sample has something to print, then in full at exactly those positions of both samples.
inline int run_batch_somatic(rv_ctx* ctx, const rv_params& P_in, const ReadBatch& batch, const std::vector<rv_region>& regs,
                             const std::vector<std::string>& genes, const std::string& refseq, int32_t ref_lo,
                             const std::string& sample, const std::string& chr, int push_flags, int halo,
                             std::string* tsv, BatchTiming* tm, std::string* err, const int64_t* read_ranges = NULL) {
  BatchTiming t;
  memset(&t, 0, sizeof t);
  int rc;
#define RV_STEP(x)                                                  \
  do {                                                              \
    rc = (x);                                                       \
    if (rc != RV_OK) {                                              \
      if (err) *err = std::string(#x) + ": " + rv_last_error(ctx);  \
      return rc;                                                    \
    }                                                               \
  } while (0)
  if (regs.size() % 2) { if (err) *err = "somatic batch needs tumor and normal regions in pairs"; return RV_ERR_ARG; }
  const size_t n = regs.size() / 2;
  rv_params P = P_in;
  P.has_bam2 = 1;
  double t0 = now_ms();
  if (push_flags & 1) RV_STEP(rv_set_reference(ctx, ref_lo, (int64_t)refseq.size(), refseq.data()));
  if (push_flags & 2) {
    rv_read_batch bv = batch.view();
    if (read_ranges) {  // {tumor lo, tumor hi, normal lo, normal hi}: only the reads of this chunk's tiles
      const int64_t lo[2] = {read_ranges[0], read_ranges[2]}, hi[2] = {read_ranges[1], read_ranges[3]};
      RV_STEP(rv_push_reads_ranges(ctx, &bv, 2, lo, hi));
      for (int k = 0; k < 2; ++k)
        if (hi[k] > lo[k]) {
          const int64_t p_lo = (int64_t)batch.reads[(size_t)lo[k]].data_off16 * 16;
          const int64_t p_hi = hi[k] < (int64_t)batch.reads.size() ? (int64_t)batch.reads[(size_t)hi[k]].data_off16 * 16
                                                                    : (int64_t)batch.pool.size();
          t.h2d_bytes += (hi[k] - lo[k]) * (int64_t)sizeof(rv_read) + (p_hi - p_lo);
        }
    } else {
      RV_STEP(rv_push_reads(ctx, &bv));
      t.h2d_bytes += (int64_t)batch.reads.size() * (int64_t)sizeof(rv_read) + (int64_t)batch.pool.size();
    }
  }
  RV_STEP(rv_set_params(ctx, &P));
  RV_STEP(rv_set_regions(ctx, regs.data(), (int32_t)regs.size()));
  double t1 = now_ms();
  RV_STEP(rv_pileup(ctx));
  RV_STEP(rv_get_pileup_stats(ctx, &t.stats));
  double t2 = now_ms();
  std::vector<int> pair_of(regs.size(), -1);
  for (size_t r = n; r < regs.size(); ++r) pair_of[r] = (int)(r - n);
  Handoff ho;
  rc = host_handoff(ctx, P, batch, regs, refseq, ref_lo, halo, &ho, &t, err, &pair_of);
  if (rc != RV_OK) return rc;
  double t5 = now_ms();
  // ---- pass A: candidate positions of either sample ----
  rv_params PA = P;
  PA.candidates_only = 2;  // positions only
  RV_STEP(rv_set_params(ctx, &PA));
  RV_STEP(rv_score(ctx));
  const rv_variant* vv;
  int64_t nv;
  RV_STEP(rv_fetch_variants(ctx, &vv, &nv));
  t.d2h_bytes += nv * (int64_t)sizeof(rv_variant);
  std::vector<std::set<int> > cand(n);
  for (int64_t i = 0; i < nv; ++i) cand[(size_t)vv[i].region % n].insert(vv[i].pos);
  std::vector<int32_t> qreg, qpos;
  for (size_t i = 0; i < n; ++i)
    for (std::set<int>::const_iterator p = cand[i].begin(); p != cand[i].end(); ++p) {
      qreg.push_back((int32_t)i); qpos.push_back(*p);
      qreg.push_back((int32_t)(i + n)); qpos.push_back(*p);
    }
  // ---- pass B: full records of both samples at those positions ----
  rv_params PB = P;
  PB.candidates_only = 0;
  RV_STEP(rv_set_params(ctx, &PB));
  RV_STEP(rv_score_positions(ctx, qreg.data(), qpos.data(), (int64_t)qreg.size()));
  RV_STEP(rv_fetch_variants(ctx, &vv, &nv));
  t.d2h_bytes += nv * (int64_t)sizeof(rv_variant);
  t.h2d_bytes += (int64_t)qreg.size() * 8;
  RV_STEP(rv_set_params(ctx, &P_in));
  {  // add_depth_by_region for both samples (somaticMode.cpp:101, :116)
    std::vector<int64_t> cs(regs.size()), cp(regs.size());
    RV_STEP(rv_cov_summary(ctx, cs.data(), cp.data()));
    for (size_t r = 0; r < regs.size(); ++r) {
      t.cov_sum[r < n ? 0 : 1] += cs[r];
      t.cov_pos[r < n ? 0 : 1] += cp[r];
    }
  }
  double t6 = now_ms();
  std::vector<int64_t> rfirst(regs.size() + 1, nv);
  {
    int64_t i = 0;
    for (size_t r = 0; r < regs.size(); ++r) {
      while (i < nv && vv[i].region < (int)r) ++i;
      rfirst[r] = i;
    }
  }
  rvk::RefView refv;
  refv.bases = refseq.data();
  refv.base_pos = ref_lo;
  refv.n = (int64_t)refseq.size();
  std::vector<std::string> rtsv(n);
  std::vector<int64_t> rlines(n, 0);
  parallel_for(n, host_threads(), [&](size_t r) {
    rvk::RefView rv = refv;
    rv.lo = regs[r].ref_lo;
    rv.hi = regs[r].ref_hi;
    // the normal sample's positions of this tile
    std::map<int, PositionVars> normal;
    std::vector<rv_variant> group;
    const size_t rn = r + n;
    for (int64_t i = rfirst[rn]; i < rfirst[rn + 1];) {
      int64_t j = i;
      while (j < rfirst[rn + 1] && vv[j].pos == vv[i].pos) ++j;
      group.assign(vv + i, vv + j);
      for (size_t k = 0; k < group.size(); ++k)
        if (group[k].key_kind == 1) group[k].key_id -= (int32_t)ho.bases[rn];
      assemble_position(P, group.data(), (int)group.size(), ho.patches[rn], rv, regs[rn].chr_len, &normal[vv[i].pos]);
      i = j;
    }
    std::string& out_s = rtsv[r];
    for (int64_t i = rfirst[r]; i < rfirst[r + 1];) {
      int64_t j = i;
      while (j < rfirst[r + 1] && vv[j].pos == vv[i].pos) ++j;
      group.assign(vv + i, vv + j);
      for (size_t k = 0; k < group.size(); ++k)
        if (group[k].key_kind == 1) group[k].key_id -= (int32_t)ho.bases[r];
      PositionVars pv;
      assemble_position(P, group.data(), (int)group.size(), ho.patches[r], rv, regs[r].chr_len, &pv);
      std::map<int, PositionVars>::iterator nit = normal.find(pv.pos);
      size_t before = out_s.size();
      output_position_somatic(P, pv, nit == normal.end() ? (PositionVars*)NULL : &nit->second, sample, genes[r], chr,
                              regs[r].start, regs[r].end, &out_s);
      for (size_t k = before; k < out_s.size(); ++k) rlines[r] += out_s[k] == '\n';
      i = j;
    }
  });
  for (size_t r = 0; r < n; ++r) {
    tsv->append(rtsv[r]);
    t.n_lines += rlines[r];
  }
  double t7 = now_ms();
  t.push_ms = t1 - t0; t.pileup_ms = t2 - t1;
  t.score_ms = t6 - t5; t.assemble_ms = t7 - t6;
  t.t_abs[0] = t0; t.t_abs[1] = t1; t.t_abs[2] = t2; t.t_abs[3] = t5; t.t_abs[4] = t6; t.t_abs[5] = t7;
  t.n_variants = nv;
  rv_last_kernel_ms(ctx, &t.pileup_kernel_ms, &t.score_kernel_ms);
  if (tm) *tm = t;
#undef RV_STEP
  return RV_OK;
}

}  // namespace rvhost
