// pileup_model.hpp — host view of one region's pileup: the dense device tables plus the sparse
// accumulators rebuilt from the device's event list in BAM order.
//
// Reference types mirrored: Variation (include/Variation.h:12-82), Sclip (include/Sclip.h:9-40), the
// per-region maps of InitialData / VariationData (include/scopedata/InitialData.h:16-21,
// VariationData.h:15-27) and CigarParser's positionToInsertionCount / positionToDeletionCount / mnp
// (include/parseCigar.h:107-116).  Dense single-base keys stay in the device layout; everything else
// lives in ordered maps so iteration is deterministic.
#pragma once
#include "../../../include/rabbitvar_b200.h"
#include "batch_loader.hpp"
#include <array>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <thread>
#include <vector>

namespace rvhost {

inline void replace_first_char(std::string& s, char c) {  // replaceFirst(s, c, "")
  size_t p = s.find(c);
  if (p != std::string::npos) s.erase(p, 1);
}

struct Variation {
  int cnt, fwd, rev;
  double sum_tp, sum_q, sum_mapq, sum_nm;
  int lo, hi;
  bool pstd, qstd;
  int pp;
  double pq;
  int extracnt;
  Variation() : cnt(0), fwd(0), rev(0), sum_tp(0), sum_q(0), sum_mapq(0), sum_nm(0), lo(0), hi(0), pstd(false),
                qstd(false), pp(0), pq(0), extracnt(0) {}
  int dir(bool d) const { return d ? rev : fwd; }
  void inc_dir(bool d) { if (d) rev++; else fwd++; }
  void add_dir(bool d, int n) { if (d) rev += n; else fwd += n; }
  void sub_dir(bool d, int n) { if (d) rev -= n; else fwd -= n; }
};

// Per-offset base histograms of a soft-clip cluster.  Built once by reduce_events_region and read-only afterwards,
// so copies of a Sclip (the realigner works on copies of the region state) share them.
struct SclipBases {
  std::map<int, std::map<char, int> > nt;
  std::map<int, std::map<char, Variation> > seq;
};
struct Sclip : Variation {
  std::shared_ptr<SclipBases> bases;
  std::string sequence;
  bool used;
  Sclip() : used(false) {}
  SclipBases& b() {
    if (!bases) bases.reset(new SclipBases());
    return *bases;
  }
  std::map<int, std::map<char, int> >& nt() { return b().nt; }
  std::map<int, std::map<char, Variation> >& seq() { return b().seq; }
  const SclipBases& cb() const {
    static const SclipBases none;
    return bases ? *bases : none;
  }
  const std::map<int, std::map<char, int> >& nt() const { return cb().nt; }
  const std::map<int, std::map<char, Variation> >& seq() const { return cb().seq; }
};

typedef std::map<std::string, Variation> KeyMap;

// addCnt, parseCigar.cpp:300-325
inline void add_cnt(Variation& v, bool dir, int tp, double q, int mapq, int nm, double goodq) {
  v.cnt++;
  v.inc_dir(dir);
  v.sum_tp += tp;
  v.sum_q += q;
  v.sum_mapq += mapq;
  v.sum_nm += nm;
  if (q >= goodq) v.hi++;
  else v.lo++;
}

// the accumulate block shared by parseCigar.cpp:888-920, :1430-1463 and addVariationForDeletion :1043-1074
inline void add_obs(Variation& v, bool dir, int tp, double q, int mapq, int nm, double goodq) {
  v.inc_dir(dir);
  v.cnt++;
  if (!v.pstd && v.pp != 0 && tp != v.pp) v.pstd = true;
  if (!v.qstd && v.pq != 0 && q != v.pq) v.qstd = true;
  v.sum_tp += tp;
  v.sum_q += q;
  v.sum_mapq += mapq;
  v.pp = tp;
  v.pq = q;
  v.sum_nm += nm;
  if (q >= goodq) v.hi++;
  else v.lo++;
}

struct RegionPileup {
  int region_idx;
  int32_t start, end;
  // dense device tables: either a full copy (tests / dumps: `dense`) or only the rows the host stage
  // asked for (rv_fetch_rows), keyed by position: 32 count words + coverage
  int32_t first_pos, n_pos;
  bool dense;
  std::vector<uint32_t> counts;  // n_pos * RV_POS_U32 when dense
  std::vector<uint32_t> cov;     // n_pos when dense
  std::map<int, std::array<uint32_t, 33> > srows;
  int64_t row_misses;            // accesses to rows that were not fetched (must stay 0)
  int max_read_len;
  RegionPileup() : region_idx(0), start(0), end(0), first_pos(0), n_pos(0), dense(true), row_misses(0), max_read_len(0) {}
  // sparse accumulators
  std::map<int, KeyMap> ni;   // nonInsertionVariants, multi-character keys (and overridden dense keys)
  std::map<int, KeyMap> ins;  // insertionVariants
  std::map<int, Sclip> sc5, sc3;
  std::map<int, std::map<std::string, int> > pins, pdel, mnp;
  // realigner bookkeeping for the write-back
  std::set<int> cov_touched;
  std::set<std::pair<int, char> > erased_dense;

  bool in_table(int pos) const { return pos >= first_pos && pos < first_pos + n_pos; }
  uint32_t* row(int pos, int allele) {
    if (dense) return counts.data() + ((size_t)(pos - first_pos) * 4 + allele) * RV_ROW_U32;
    std::map<int, std::array<uint32_t, 33> >::iterator it = srows.find(pos);
    if (it == srows.end()) {
      row_misses++;
      std::array<uint32_t, 33> z;
      z.fill(0);
      it = srows.insert(std::make_pair(pos, z)).first;
    }
    return it->second.data() + allele * RV_ROW_U32;
  }
  uint32_t& cov_at(int pos) {
    if (dense) return cov[(size_t)(pos - first_pos)];
    return row(pos, 0)[32];
  }
  static bool row_exists(const uint32_t* r) {
    uint32_t o = 0;
    for (int k = 0; k < RV_ROW_U32; ++k) o |= r[k];
    return o != 0;
  }
  static Variation row_to_variation(const uint32_t* r) {
    Variation v;
    v.fwd = (int)r[RV_F_FWD];
    v.rev = (int)r[RV_F_REV];
    v.cnt = v.fwd + v.rev;
    v.hi = (int)r[RV_F_HI];
    v.lo = v.cnt - v.hi;
    v.sum_tp = (double)(int)r[RV_F_SUM_TP];
    v.sum_q = (double)(int)r[RV_F_SUM_Q];
    v.sum_mapq = (double)(int)r[RV_F_SUM_MAPQ];
    v.sum_nm = (double)(int)r[RV_F_SUM_NM];
    v.pstd = (r[RV_F_STD] >> 24) & 1;
    v.qstd = (r[RV_F_STD] >> 25) & 1;
    return v;
  }
};

inline int allele_index(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }

// Rebuilds the sparse accumulators of ONE region from its slice of the BAM-ordered event list.
inline void reduce_events_region(const rv_event* ev, int64_t n, const ReadBatch& batch, double goodq, RegionPileup& R) {
  Variation* last_ins = NULL;
  uint32_t last_ins_read = 0xffffffffu;
  for (int64_t i = 0; i < n; ++i) {
    const rv_event& e = ev[i];
    // a key that did not fit RV_EVENT_KEY_MAX is not scored under a wrong name: the device counted the read in
    // n_unsupported, the observation is dropped here
    if (e.flags & RV_EVF_KEY_TRUNC) continue;
    const bool dir = e.dir != 0;
    const double q = e.qsum / (double)e.qcnt;
    std::string key(e.key, e.keylen);
    switch (e.kind) {
      case RV_EV_NI: {
        Variation& v = R.ni[e.pos][key];
        if (e.flags & RV_EVF_MNP) R.mnp[e.pos][key]++;
        add_obs(v, dir, e.tp, q, e.mapq, e.nm, goodq);
        if (e.flags & RV_EVF_PDEL) R.pdel[e.pos][key]++;
        break;
      }
      case RV_EV_IN: {
        R.pins[e.pos][key]++;
        Variation& v = R.ins[e.pos][key];
        add_obs(v, dir, e.tp, q, e.mapq, e.nm, goodq);
        last_ins = &v;
        last_ins_read = e.read_idx;
        break;
      }
      case RV_EV_SC5: {
        Sclip& s = R.sc5[e.pos];
        const int m = e.aux0, nhi = e.aux1;
        for (int si = m - 1; m - si <= nhi; si--) {
          char ch = batch.base(e.read_idx, si);
          int idx = m - 1 - si;
          s.nt()[idx][ch]++;
          add_cnt(s.seq()[idx][ch], dir, si - (m - nhi), batch.qual(e.read_idx)[si], e.mapq, e.nm, goodq);
        }
        add_cnt(s, dir, m, q, e.mapq, e.nm, goodq);
        break;
      }
      case RV_EV_SC3: {
        Sclip& s = R.sc3[e.pos];
        const int m = e.aux0, nhi = e.aux1, rp = e.aux2;
        for (int si = 0; si < nhi; si++) {
          char ch = batch.base(e.read_idx, rp + si);
          s.nt()[si][ch]++;
          add_cnt(s.seq()[si][ch], dir, nhi - si, batch.qual(e.read_idx)[rp + si], e.mapq, e.nm, goodq);
        }
        add_cnt(s, dir, m, q, e.mapq, e.nm, goodq);
        break;
      }
      case RV_EV_TTREF: {
        // parseCigar.cpp:1497-1515 — one extra observation on the reference allele; the dense row is
        // moved into the sparse map so the pstd/qstd assignment can be represented
        int al = allele_index(key[0]);
        if (al < 0 || !R.in_table(e.pos)) break;
        KeyMap& km = R.ni[e.pos];
        if (!km.count(key)) km[key] = RegionPileup::row_to_variation(R.row(e.pos, al));
        Variation& t = km[key];
        t.inc_dir(dir);
        t.cnt++;
        if (last_ins && last_ins_read == e.read_idx) { t.pstd = last_ins->pstd; t.qstd = last_ins->qstd; }
        t.sum_tp += e.tp;
        t.sum_q += q;
        t.sum_mapq += e.mapq;
        t.pp = e.tp;
        t.pq = q;
        t.sum_nm += e.nm;
        break;
      }
      default: break;
    }
  }
}

// Simple static-partition parallel loop over [0, n) on up to `threads` host threads.
template <class F>
inline void parallel_for(size_t n, int threads, F f) {
  if (threads <= 1 || n <= 1) {
    for (size_t i = 0; i < n; ++i) f(i);
    return;
  }
  std::vector<std::thread> th;
  size_t nt = std::min<size_t>((size_t)threads, n);
  for (size_t t = 0; t < nt; ++t)
    th.emplace_back([=]() {
      for (size_t i = t; i < n; i += nt) f(i);
    });
  for (size_t t = 0; t < th.size(); ++t) th[t].join();
}
inline int& host_threads_override() {  // per calling thread: workers of a pipeline share the cores
  static thread_local int v = 0;
  return v;
}
inline int host_threads() {
  if (host_threads_override() > 0) return host_threads_override();
  const char* e = getenv("RV_HOST_THREADS");
  if (e && atoi(e) > 0) return atoi(e);
  unsigned hc = std::thread::hardware_concurrency();
  return hc == 0 ? 4 : (int)std::min(hc, 32u);
}

// Rebuilds the sparse accumulators of every region (events are sorted by region, then BAM order).
inline void reduce_events(const rv_event* ev, int64_t n, const ReadBatch& batch, double goodq,
                          std::vector<RegionPileup>& regions) {
  std::vector<int64_t> first(regions.size() + 1, n);
  int64_t i = 0;
  for (size_t r = 0; r < regions.size(); ++r) {
    while (i < n && ev[i].region < (int)r) ++i;
    first[r] = i;
  }
  first[regions.size()] = n;
  for (size_t r = regions.size(); r-- > 0;)
    if (first[r] > first[r + 1]) first[r] = first[r + 1];
  parallel_for(regions.size(), host_threads(), [&](size_t r) {
    int64_t a = first[r], z = first[r + 1];
    while (z > a && ev[z - 1].region != (int)r) --z;
    reduce_events_region(ev + a, z - a, batch, goodq, regions[r]);
  });
}

}  // namespace rvhost
