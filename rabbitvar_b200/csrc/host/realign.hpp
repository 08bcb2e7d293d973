// realign.hpp — host-side stage between pileup and scoring (VariationRealigner::process, reference
// src/VariationRealigner.cpp:135-163) on the RegionPileup model, and the write-back of its result as
// patch entries for rv_apply_patch.  north_star keeps this stage on the host.
//
// Implemented: adjustMNP (:334-465) with findconseq (include/VariationUtils.h:403-509) and ismatchref
// (:2004-2024); realignIndels (:467-1354) = realigndel, realignins, realignlgins30 with findMM3/findMM5,
// ismatch, find35match, noPassingReads, adjCnt/adjRefCnt/adjRefFactor.  Structural-variant keys (<dup..>, <inv..>)
// never reach this path (the SV code of the reference is commented out).
#pragma once
#include <atomic>
#include "pileup_model.hpp"
#include "../kernels/rv_core.cuh"
#include <string.h>
#include <algorithm>
#include <set>

namespace rvhost {

// ---- unified view of nonInsertionVariants[pos] over the dense rows and the sparse map ------------
// A dense single-base key that the realigner needs to touch is first "promoted": copied into the sparse
// map and cleared in the dense table, so every key lives in exactly one place.
struct NiView {
  RegionPileup& R;
  std::set<int> cov_touched;
  std::set<std::pair<int, char> > erased_dense;  // dense keys removed from the map (tombstones for the device)
  explicit NiView(RegionPileup& r) : R(r) {}

  void promote_all(int pos) {
    if (!R.in_table(pos)) return;
    for (int a = 0; a < 4; ++a) {
      uint32_t* row = R.row(pos, a);
      if (RegionPileup::row_exists(row)) {
        std::string k(1, "ACGT"[a]);
        KeyMap& km = R.ni[pos];
        if (!km.count(k)) km[k] = RegionPileup::row_to_variation(row);
        memset(row, 0, sizeof(uint32_t) * RV_ROW_U32);
        promoted.insert(std::make_pair(pos, "ACGT"[a]));
      }
    }
  }
  bool has_pos(int pos) {
    if (R.ni.count(pos) && !R.ni[pos].empty()) return true;
    if (!R.in_table(pos)) return false;
    for (int a = 0; a < 4; ++a) if (RegionPileup::row_exists(R.row(pos, a))) return true;
    return false;
  }
  // nonInsertionVariants->count(pos): the position map entry exists once any key was ever created
  bool count_pos(int pos) { return has_pos(pos) || R.ni.count(pos); }
  KeyMap& keys(int pos) { promote_all(pos); return R.ni[pos]; }
  Variation* find(int pos, const std::string& key) {
    promote_all(pos);
    std::map<int, KeyMap>::iterator it = R.ni.find(pos);
    if (it == R.ni.end()) return NULL;
    KeyMap::iterator k = it->second.find(key);
    return k == it->second.end() ? NULL : &k->second;
  }
  void erase(int pos, const std::string& key) {
    promote_all(pos);
    std::map<int, KeyMap>::iterator it = R.ni.find(pos);
    if (it == R.ni.end()) return;
    it->second.erase(key);
    if (key.size() == 1 && promoted.count(std::make_pair(pos, key[0]))) erased_dense.insert(std::make_pair(pos, key[0]));
  }
  void add_cov(int pos, int n) {
    if (!R.in_table(pos)) return;
    R.cov_at(pos) += (uint32_t)n;
    cov_touched.insert(pos);
  }
  std::set<std::pair<int, char> > promoted;
};

// adjCnt without reference variant, include/VariationUtils.h:283-296
inline void adj_cnt(Variation& to, const Variation& v) {
  to.cnt += v.cnt;
  to.extracnt += v.cnt;
  to.hi += v.hi;
  to.lo += v.lo;
  to.sum_tp += v.sum_tp;
  to.sum_q += v.sum_q;
  to.sum_mapq += v.sum_mapq;
  to.sum_nm += v.sum_nm;
  to.pstd = true;
  to.qstd = true;
  to.rev += v.rev;
  to.fwd += v.fwd;
}

inline bool is_low_complex(const std::string& seq) {  // islowcomplexseq, VariationUtils.h:360-393
  int len = (int)seq.size();
  if (len == 0) return true;
  int ntcnt = 0;
  const char order[4] = {'A', 'T', 'G', 'C'};
  for (int k = 0; k < 4; ++k) {
    int c = (int)std::count(seq.begin(), seq.end(), order[k]);
    if (c > 0) ntcnt++;
    if (c / (double)len > 0.75) return true;
  }
  return ntcnt < 3;
}

// findconseq, include/VariationUtils.h:403-509.  The reference walks each per-offset base histogram in
// robin_hood hash order; here bases are visited in `order_hint` order (parity notes).
inline std::string find_conseq(Sclip& sc) {
  if (!sc.sequence.empty()) return sc.sequence;
  int total = 0, match = 0;
  std::string seqq;
  bool flag = false;
  for (std::map<int, std::map<char, int> >::iterator nve = sc.nt().begin(); nve != sc.nt().end(); ++nve) {
    int pis = nve->first;
    int maxCount = 0;
    double maxQuality = 0;
    char chosen = 0;
    int totalCount = 0;
    // The reference walks a robin_hood::unordered_map<char, int> here and its choice depends on the visiting order
    // (a later base with fewer reads but a larger quality sum replaces an earlier one).  For the keys that can
    // occur the flat map's slot order is fixed, whatever the insertion order: N, A, T, G, C (probed against the
    // reference's own robin_hood.h 3.4.3 for every subset and insertion order: oracle/robin_hood_order.cpp).
    // Any other IUPAC letter follows in ASCII order.
    std::vector<std::pair<char, int> > ordered;
    {
      static const char RH_ORDER[] = "NATGC";
      for (const char* o = RH_ORDER; *o; ++o) {
        std::map<char, int>::iterator f = nve->second.find(*o);
        if (f != nve->second.end()) ordered.push_back(*f);
      }
      for (std::map<char, int>::iterator ent = nve->second.begin(); ent != nve->second.end(); ++ent)
        if (!strchr(RH_ORDER, ent->first)) ordered.push_back(*ent);
    }
    for (size_t oi = 0; oi < ordered.size(); ++oi) {
      char cb = ordered[oi].first;
      int cc = ordered[oi].second;
      totalCount += cc;
      bool hasq = sc.seq().count(pis) && sc.seq()[pis].count(cb);
      if (cc > maxCount || (hasq && sc.seq()[pis][cb].sum_q > maxQuality)) {
        maxCount = cc;
        chosen = cb;
        maxQuality = sc.seq()[pis][cb].sum_q;
      }
    }
    if (pis == 3 && sc.nt().size() >= 6 && totalCount / (double)sc.cnt < 0.2 && totalCount <= 2) break;
    if ((totalCount - maxCount > 2 || maxCount <= totalCount - maxCount) && maxCount / (double)totalCount < 0.8) {
      if (flag) break;
      flag = true;
    }
    total += totalCount;
    match += maxCount;
    if (chosen != 0) seqq += chosen;
  }
  std::string SEQ;
  int ntSize = (int)sc.nt().size();
  if (total != 0 && match / (double)total > 0.9 && seqq.size() / 1.5 > ntSize - (double)seqq.size() &&
      (seqq.size() / (double)ntSize > 0.8 || ntSize - (int)seqq.size() < 10 || seqq.size() > 25))
    SEQ = seqq;
  else
    SEQ = " ";
  if (!SEQ.empty() && SEQ.size() > 12) {  // CONF_SEED_2
    bool a7 = SEQ.size() >= 8 && SEQ.compare(1, 7, "AAAAAAA") == 0;
    bool t7 = SEQ.size() >= 8 && SEQ.compare(1, 7, "TTTTTTT") == 0;
    if (a7 || t7) sc.used = true;
    if (is_low_complex(SEQ)) sc.used = true;
  }
  sc.sequence = SEQ;
  return SEQ;
}

// ismatchref, VariationRealigner.cpp:2004-2024
inline bool is_match_ref(const std::string& seq, const rvk::RefView& ref, int position, int dir, int MM = 3) {
  int mm = 0;
  for (int n = 0; n < (int)seq.size(); n++) {
    if (!ref.has(position + dir * n)) return false;
    int idx = dir == 1 ? n : dir * n - 1;
    char c;
    if (idx < 0) { int i = (int)seq.size() + idx; c = i < 0 ? (char)-1 : seq[i]; }
    else c = seq[idx];
    if (c != ref.at(position + dir * n)) mm++;
  }
  return mm <= MM && mm / (double)seq.size() < 0.15;
}

struct PosDesc { int position; std::string desc; int count; };
inline bool cmp_tmp(const PosDesc& a, const PosDesc& b) {  // CMP_tmp, VariationRealigner.cpp:44-58
  if (a.count != b.count) return a.count > b.count;
  if (a.position != b.position) return a.position < b.position;
  return a.desc.compare(b.desc) > 0;
}
inline std::vector<PosDesc> fill_and_sort(const std::map<int, std::map<std::string, int> >& m) {
  std::vector<PosDesc> t;
  for (std::map<int, std::map<std::string, int> >::const_iterator p = m.begin(); p != m.end(); ++p)
    for (std::map<std::string, int>::const_iterator k = p->second.begin(); k != p->second.end(); ++k) {
      PosDesc d;
      d.position = p->first; d.desc = k->first; d.count = k->second;
      t.push_back(d);
    }
  std::sort(t.begin(), t.end(), cmp_tmp);
  return t;
}

// adjustMNP, VariationRealigner.cpp:334-465
inline void adjust_mnp(NiView& V, const rvk::RefView& ref) {
  RegionPileup& R = V.R;
  std::vector<PosDesc> tmp = fill_and_sort(R.mnp);
  for (size_t ti = 0; ti < tmp.size(); ++ti) {
    const int position = tmp[ti].position;
    const std::string vn = tmp[ti].desc;
    if (!V.count_pos(position)) continue;
    // the reference works on a COPY of the position's key->Variation* map: erasing from it does not
    // touch the real map, but the Variation objects are shared (:348, :378)
    KeyMap& real = V.keys(position);
    std::set<std::string> erased_in_copy;
    if (!real.count(vn)) continue;
    Variation* vref = &real[vn];
    std::string mnt = vn;
    replace_first_char(mnt, '&');
    for (int i = 0; i < (int)mnt.size() - 1; i++) {
      std::string left = mnt.substr(0, i + 1);
      if (left.size() > 1) left.insert(1, "&");
      std::string right = mnt.substr(i + 1);
      if (right.size() > 1) right.insert(1, "&");
      if (real.count(left) && !erased_in_copy.count(left)) {
        Variation* tref = &real[left];
        if (tref->cnt <= 0) continue;
        if (tref->cnt < vref->cnt && tref->sum_tp / tref->cnt <= i + 1) {
          adj_cnt(*vref, *tref);
          erased_in_copy.insert(left);
        }
      }
      if (V.count_pos(position + i + 1)) {
        Variation* tref = V.find(position + i + 1, right);
        if (tref) {
          if (tref->cnt < 0) continue;
          if (tref->cnt < vref->cnt) {
            adj_cnt(*vref, *tref);
            V.add_cov(position, tref->cnt);
            V.erase(position + i + 1, right);
          }
        }
      }
    }
    if (R.sc3.count(position)) {
      Sclip& sc3v = R.sc3[position];
      if (!sc3v.used) {
        const std::string seq = find_conseq(sc3v);
        if (seq.substr(0, mnt.size()) == mnt) {
          if (seq.size() == mnt.size() || is_match_ref(seq.substr(mnt.size()), ref, position + (int)mnt.size(), 1)) {
            adj_cnt(V.keys(position)[vn], sc3v);
            V.add_cov(position, sc3v.cnt);
            sc3v.used = true;
          }
        }
      }
    }
    if (R.sc5.count(position + (int)mnt.size())) {
      Sclip& sc5v = R.sc5[position + (int)mnt.size()];
      if (!sc5v.used) {
        std::string seq = find_conseq(sc5v);
        if (seq != " " && seq.size() >= mnt.size()) {
          std::reverse(seq.begin(), seq.end());
          if (seq.substr(seq.size() - mnt.size(), mnt.size()) == mnt) {
            if (seq.size() == mnt.size() ||
                is_match_ref(seq.substr(0, seq.size() - mnt.size()), ref, position - 1, -1)) {
              adj_cnt(V.keys(position)[vn], sc5v);
              V.add_cov(position, sc5v.cnt);
              sc5v.used = true;
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// realignIndels, VariationRealigner.cpp:467-1354: realigndel, realignins, realignlgins30 and their helpers.
// The reference wraps every iteration of its loops in try/catch(...): a failing map .at() or substr abandons the
// rest of that iteration with its side effects kept.  RealignAbort models those exceptions.
// ------------------------------------------------------------------------------------------------
struct RealignAbort {};

inline std::string join_ref_at(const rvk::RefView& ref, int a, int b) {  // joinRef uses .at(): throws outside the window
  std::string s;
  for (int i = a; i <= b; ++i) {
    if (!ref.has(i)) throw RealignAbort();
    s.push_back(ref.at(i));
  }
  return s;
}
inline std::string join_ref_double(const rvk::RefView& ref, int a, double to) {  // joinRef_double: i < to
  std::string s;
  for (int i = a; i < to; ++i) {
    if (!ref.has(i)) throw RealignAbort();
    s.push_back(ref.at(i));
  }
  return s;
}
inline char char_at(const std::string& s, int index) {  // charAt, util.h:171-179
  if (index < 0) {
    int i = (int)s.size() + index;
    if (i < 0) return (char)-1;
    return s[(size_t)i];
  }
  return (size_t)index < s.size() ? s[(size_t)index] : (char)0;
}
inline std::string vc_substr1(const std::string& s, int idx) {  // vc_substr(str, idx), util.h:27-34
  if (idx >= 0) {
    if ((size_t)idx > s.size()) throw RealignAbort();
    return s.substr((size_t)idx);
  }
  // `str.length() + idx < 0` is an unsigned comparison in the reference and never true
  if ((size_t)(-idx) > s.size()) throw RealignAbort();
  return s.substr(s.size() - (size_t)(-idx));
}
inline std::string vc_substr2(const std::string& s, int begin, int len) {  // vc_substr(str, begin, len), util.h:44-57
  if (begin < 0) begin = (int)s.size() + begin;
  if (len == 0) return "";
  if (begin < 0 || (size_t)begin > s.size()) throw RealignAbort();
  if (len > 0) return s.substr((size_t)begin, (size_t)len);
  len = (int)s.size() + len - begin;
  if (len < 0) return "";
  return s.substr((size_t)begin, (size_t)len);
}
inline std::string strip_hash_caret(const std::string& s) {  // regex_replace(s, "#|\\^", "")
  std::string o;
  for (size_t i = 0; i < s.size(); ++i) if (s[i] != '#' && s[i] != '^') o.push_back(s[i]);
  return o;
}
inline bool is_atgnc(char c) { return c == 'A' || c == 'T' || c == 'G' || c == 'N' || c == 'C'; }
inline bool is_atgc(char c) { return c == 'A' || c == 'T' || c == 'G' || c == 'C'; }
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }

// correctCnt, VariationUtils.h:258-275
inline void correct_cnt(Variation& v) {
  if (v.cnt < 0) v.cnt = 0;
  if (v.hi < 0) v.hi = 0;
  if (v.lo < 0) v.lo = 0;
  if (v.sum_tp < 0) v.sum_tp = 0;
  if (v.sum_q < 0) v.sum_q = 0;
  if (v.sum_mapq < 0) v.sum_mapq = 0;
  if (v.rev < 0) v.rev = 0;
  if (v.fwd < 0) v.fwd = 0;
}
// adjCnt with reference variant, VariationUtils.h:286-313 (`v` and `ref` may be the same object)
inline void adj_cnt3(Variation& to, Variation& v, Variation* ref) {
  adj_cnt(to, v);
  if (!ref) return;
  const Variation c = v;
  ref->cnt -= c.cnt;
  ref->hi -= c.hi;
  ref->lo -= c.lo;
  ref->sum_tp -= c.sum_tp;
  ref->sum_q -= c.sum_q;
  ref->sum_mapq -= c.sum_mapq;
  ref->sum_nm -= c.sum_nm;
  ref->rev -= c.rev;
  ref->fwd -= c.fwd;
  correct_cnt(*ref);
}
// adjRefCnt, VariationRealigner.cpp:1753-1787
inline void adj_ref_cnt(const Variation& tv, Variation* ref, int len) {
  if (!ref) return;
  double f = tv.sum_tp != 0 ? (tv.sum_tp / (double)tv.cnt - len + 1) / (tv.sum_tp / (double)tv.cnt) : 0;
  if (f < 0) return;
  if (f > 1) f = 1;
  ref->cnt -= (int)(f * tv.cnt);
  ref->hi -= (int)(f * tv.hi);
  ref->lo -= (int)(f * tv.lo);
  ref->sum_tp -= f * tv.sum_tp;
  ref->sum_q -= f * tv.sum_q;
  ref->sum_mapq -= f * tv.sum_mapq;
  ref->sum_nm -= f * tv.sum_nm;
  ref->rev -= (int)(f * tv.rev);
  ref->fwd -= (int)(f * tv.fwd);
  correct_cnt(*ref);
}
// adjRefFactor, VariationRealigner.cpp:1794-1822
inline void adj_ref_factor(Variation* ref, double factor_f) {
  if (!ref) return;
  if (factor_f > 1) factor_f = 1;
  if (factor_f < -1) return;
  int old = ref->cnt;
  ref->cnt -= (int)(factor_f * ref->cnt);
  ref->hi -= (int)(factor_f * ref->hi);
  ref->lo -= (int)(factor_f * ref->lo);
  int d = ref->cnt - old;
  double factor_cnt = old != 0 ? (d < 0 ? -d : d) / (double)old : 1;
  ref->sum_tp -= ref->sum_tp * factor_f * factor_cnt;
  ref->sum_q -= ref->sum_q * factor_f * factor_cnt;
  ref->sum_mapq -= ref->sum_mapq * factor_f * factor_cnt;
  ref->sum_nm -= factor_f * ref->sum_nm;
  ref->fwd -= (int)(factor_f * ref->fwd);
  ref->rev -= (int)(factor_f * ref->rev);
  correct_cnt(*ref);
}

struct Mismatch { std::string seq; int pos; int end; };
struct MismatchResult {
  std::vector<Mismatch> mm;
  std::vector<int> scp;
  int nm, misp;
  std::string misnt;
};

// ismatch, VariationRealigner.cpp:1507-1540
inline bool is_match(const std::string& seq1, const std::string& seq2_in, int dir, int MM = 3) {
  const std::string seq2 = strip_hash_caret(seq2_in);
  int mm = 0;
  for (size_t n = 0; n < seq1.size() && n < seq2.size(); n++) {
    const std::string c2 = vc_substr2(seq2, dir * (int)n - (dir == -1 ? 1 : 0), 1);
    if (seq1[n] != (c2.empty() ? (char)0 : c2[0])) mm++;
  }
  return mm <= MM && mm / (double)seq1.size() < 0.15;
}

// Everything the realigner works on.
struct Realigner {
  const rv_params& P;
  RegionPileup& R;
  NiView& V;
  const rvk::RefView& ref;
  int chr_len;
  const ReadBatch* batch;  // for noPassingReads; NULL = no BAM access (then the test is skipped like bams == NULL)
  int64_t read_lo, read_hi;
  Realigner(const rv_params& p, RegionPileup& r, NiView& v, const rvk::RefView& rf, int cl, const ReadBatch* b, int64_t lo,
            int64_t hi)
      : P(p), R(r), V(v), ref(rf), chr_len(cl), batch(b), read_lo(lo), read_hi(hi) {}

  char ref_char(int p) const { return ref.at(p); }  // ref[p]: '\0' when the position is not loaded
  Variation* ref_var(int p) {                      // getVariationMaybe(nonInsertionVariants, p, ref[p])
    char c = ref_char(p);
    if (c == 0) return NULL;
    if (!V.count_pos(p)) return NULL;
    return V.find(p, std::string(1, c));
  }
  void erase_ni(int pos, const std::string& key, bool drop_empty_position) {
    V.erase(pos, key);
    if (drop_empty_position) {
      std::map<int, KeyMap>::iterator it = R.ni.find(pos);
      if (it != R.ni.end() && it->second.empty()) R.ni.erase(it);
    }
  }

  // findMM5, VariationRealigner.cpp:1857-1909
  MismatchResult find_mm5(int position, const std::string& wupseq) {
    const std::string seq = strip_hash_caret(wupseq);
    const int longmm = 3;
    MismatchResult r;
    int n = 0, mn = 0, mcnt = 0;
    std::string str;
    while (rvk::has_ne(ref, position - n, char_at(seq, -1 - n)) && mcnt < longmm) {
      str.insert(str.begin(), char_at(seq, -1 - n));
      Mismatch m; m.seq = str; m.pos = position - n; m.end = 5;
      r.mm.push_back(m);
      n++;
      mcnt++;
    }
    r.scp.push_back(position + 1);
    int misp = 0;
    char misnt = 0;
    if (str.size() == 1) {
      while (rvk::has_eq(ref, position - n, char_at(seq, -1 - n))) {
        n++;
        if (n != 0) mn++;
      }
      if (mn > 1) {
        int n2 = 0;
        while (-1 - n - 1 - n2 >= 0 && rvk::has_eq(ref, position - n - 1 - n2, char_at(seq, -1 - n - 1 - n2))) n2++;
        if (n2 > 2) {
          r.scp.push_back(position - n - n2);
          misp = position - n;
          misnt = char_at(seq, -1 - n);
          if (R.sc5.count(position - n - n2)) R.sc5[position - n - n2].used = true;
          mn += n2;
        } else {
          r.scp.push_back(position - n);
          if (R.sc5.count(position - n)) R.sc5[position - n].used = true;
        }
      }
    }
    r.nm = mn;
    r.misp = misp;
    r.misnt = misnt == 0 ? std::string() : std::string(1, misnt);
    return r;
  }

  // findMM3, VariationRealigner.cpp:1919-1975
  MismatchResult find_mm3(int p, const std::string& sanpseq) {
    const std::string seq = strip_hash_caret(sanpseq);
    const int longmm = 3;
    const int len = (int)seq.size();
    MismatchResult r;
    int n = 0, mn = 0, mcnt = 0;
    std::string str;
    while (n < len && ref.has(p + n) && ref.at(p + n) == seq[(size_t)n]) n++;
    r.scp.push_back(p + n);
    const int Tbp = p + n;
    while (mcnt <= longmm && n < len && ref.at(p + n) != seq[(size_t)n]) {
      str += seq[(size_t)n];
      Mismatch m; m.seq = str; m.pos = Tbp; m.end = 3;
      r.mm.push_back(m);
      n++;
      mcnt++;
    }
    int misp = 0;
    char misnt = 0;
    if (str.size() == 1) {
      while (n < len && rvk::has_eq(ref, p + n, seq[(size_t)n])) {
        n++;
        if (n != 0) mn++;
      }
      if (mn > 1) {
        int n2 = 0;
        while (n + n2 + 1 < len && rvk::has_eq(ref, p + n + 1 + n2, seq[(size_t)(n + n2 + 1)])) n2++;
        if (n2 > 2 && n + n2 + 1 < len) {
          r.scp.push_back(p + n + n2);
          misp = p + n;
          misnt = seq[(size_t)n];
          if (R.sc3.count(p + n + n2)) R.sc3[p + n + n2].used = true;
          mn += n2;
        } else {
          r.scp.push_back(p + n);
          if (R.sc3.count(p + n)) R.sc3[p + n].used = true;
        }
      }
    }
    r.nm = mn;
    r.misp = misp;
    r.misnt = misnt == 0 ? std::string() : std::string(1, misnt);
    return r;
  }

  // noPassingReads, VariationRealigner.cpp:1447-1500: every record of the BAM overlapping chr:start-end
  bool no_passing_reads(int start, int end) {
    int cnt = 0, midcnt = 0;
    const int dlen = end - start;
    const uint32_t dlenqr = ((uint32_t)dlen << 4) | 2u;
    // reads are sorted by start: only those starting in [start - longest reference span, end] can overlap.  The
    // region's slice [read_lo, read_hi) ends at the region's own end; an indel beyond it (the reads that carry it overlap
    // the region, the reads that pass it need not) is looked up in the rest of the sorted run the slice belongs to —
    // the neighbouring tiles' reads of the same sample and the margin the loader reads beyond the span.
    const int64_t n_all = (int64_t)batch->reads.size();
    int64_t i0 = read_lo, z = read_hi;
    const int64_t want = (int64_t)start - 1 - batch->max_ref_span;
    while (i0 < z) { const int64_t m = (i0 + z) / 2; if ((int64_t)batch->reads[(size_t)m].pos - 1 < want) i0 = m + 1; else z = m; }
    if (i0 == read_lo && read_lo < n_all)
      while (i0 > 0 && batch->reads[(size_t)i0 - 1].pos <= batch->reads[(size_t)i0].pos && (int64_t)batch->reads[(size_t)i0 - 1].pos - 1 >= want) --i0;
    for (int64_t i = i0; i < n_all; ++i) {
      const rv_read& rd = batch->reads[(size_t)i];
      if (i >= read_hi && i > 0 && rd.pos < batch->reads[(size_t)i - 1].pos) break;  // the next span / sample begins
      if (rd.pos - 1 >= end) break;
      if (!(rd.pos - 1 < end && rd.end_pos > start - 1)) continue;  // sam_itr_querys("chr:start-end")
      const uint32_t* cg = batch->cigar((size_t)i);
      bool has = false;
      int aligned = 0;
      for (int k = 0; k < rd.n_cigar; ++k) {
        if (cg[k] == dlenqr) has = true;
        const int op = (int)(cg[k] & 0xf);
        if (op == 0 || op == 2) aligned += (int)(cg[k] >> 4);
      }
      if (has) continue;
      const int read_start = rd.pos, read_end = read_start + aligned;
      if (read_end > end + 2 && read_start < start - 2) cnt++;
      if (read_start < start - 2 && read_end > start && read_end < end) midcnt++;
    }
    return cnt <= 0 && midcnt + 1 > 0;
  }

  // the loop over mismatching bases next to an indel shared by realigndel (:582-647) and realignins (:918-964)
  // realigndel, VariationRealigner.cpp:489-778
  void realign_del(bool bams, const std::map<int, std::map<std::string, int> >& pdel) {
    std::vector<PosDesc> tmp = fill_and_sort(pdel);
    for (size_t ti = 0; ti < tmp.size(); ++ti) {
      try {
        const int p = tmp[ti].position;
        const std::string vn = tmp[ti].desc;
        const int dcnt = tmp[ti].count;
        Variation* vref = &V.keys(p)[vn];  // getVariation
        int dellen = 0;
        if (vn.size() > 1 && vn[0] == '-' && is_digit(vn[1])) dellen = atoi(vn.c_str() + 1);  // ^-(\d+).*
        {  // \^(\d+)$
          size_t e = vn.size();
          size_t k = e;
          while (k > 0 && is_digit(vn[k - 1])) --k;
          if (k < e && k > 0 && vn[k - 1] == '^') dellen += atoi(vn.c_str() + k);
        }
        std::string extrains, extra, inv5, inv3;
        if (vn.find('<') != std::string::npos) {
          // ^-\d+\^([ATGNC]+)<...\d+>([ATGNC]+)$ — structural-variant keys are never produced by this path
          throw RealignAbort();
        } else if (vn.size() > 1 && vn[0] == '-' && is_digit(vn[1])) {  // ^-\d+(.*)
          size_t k = 1;
          while (k < vn.size() && is_digit(vn[k])) ++k;
          for (size_t j = k; j < vn.size(); ++j)
            if (vn[j] != '^' && vn[j] != '&' && vn[j] != '#') extra.push_back(vn[j]);
          for (size_t j = 0; j + 1 < vn.size(); ++j)  // \^([ATGNC]+): first '^' followed by such a run
            if (vn[j] == '^' && is_atgnc(vn[j + 1])) {
              size_t z = j + 1;
              while (z < vn.size() && is_atgnc(vn[z])) ++z;
              extrains = vn.substr(j + 1, z - j - 1);
              break;
            }
        }
        const int wustart = (p - 200) > 1 ? (p - 200) : 1;
        std::string wupseq = join_ref_at(ref, wustart, p - 1) + extra;
        if (!inv3.empty()) wupseq = inv3;
        const int sanend = (chr_len > 0 && p + 200 > chr_len) ? chr_len : p + 200;
        const int tail = p + dellen + (int)extra.size() - (int)extrains.size();
        std::string sanpseq = extra + join_ref_at(ref, tail, sanend);
        if (!inv5.empty()) sanpseq = inv5;
        MismatchResult r3 = find_mm3(p, sanpseq);
        MismatchResult r5 = find_mm5(tail - 1, wupseq);
        const int nm3 = r3.nm, nm5 = r5.nm;
        std::vector<Mismatch> mmm(r3.mm);
        mmm.insert(mmm.end(), r5.mm.begin(), r5.mm.end());
        for (size_t mi = 0; mi < mmm.size(); ++mi) {
          std::string mm = mmm[mi].seq;
          const int mp = mmm[mi].pos, me = mmm[mi].end;
          if (mm.size() > 1) mm.insert(1, "&");
          if (!V.count_pos(mp)) continue;
          Variation* tv = V.find(mp, mm);
          if (!tv) continue;
          if (tv->cnt == 0) continue;
          if (tv->sum_q / tv->cnt < P.goodq) continue;
          if (tv->sum_tp / tv->cnt > (me == 3 ? nm3 + 4 : nm5 + 4)) continue;
          if (tv->cnt >= dcnt + dellen || tv->cnt / dcnt >= 8) continue;
          if (mp > p && me == 5) {  // adjust ref cnt so that AF won't > 1
            double f = tv->sum_tp != 0 ? (mp - p) / (tv->sum_tp / (double)tv->cnt) : 1;
            if (f > 1) f = 1;
            V.add_cov(p, (int)(tv->cnt * f));
            adj_ref_cnt(*tv, ref_var(p), dellen);
          }
          Variation* lref = NULL;
          if (mp > p && me == 3 && V.count_pos(p)) lref = V.find(p, std::string(1, ref_char(p)));
          if (lref) adj_cnt3(*vref, *tv, lref);
          else adj_cnt(*vref, *tv);
          erase_ni(mp, mm, true);
        }
        if (r3.misp != 0 && r3.mm.size() == 1 && V.count_pos(r3.misp)) {
          Variation* t = V.find(r3.misp, r3.misnt);
          if (t && t->cnt < dcnt) erase_ni(r3.misp, r3.misnt, false);
        }
        if (r5.misp != 0 && r5.mm.size() == 1 && V.count_pos(r5.misp)) {
          Variation* t = V.find(r5.misp, r5.misnt);
          if (t && t->cnt < dcnt) erase_ni(r5.misp, r5.misnt, false);
        }
        for (size_t k = 0; k < r5.scp.size(); ++k) {
          const int sc5pp = r5.scp[k];
          std::map<int, Sclip>::iterator it = R.sc5.find(sc5pp);
          if (it == R.sc5.end() || it->second.used) continue;
          Sclip& tv = it->second;
          const std::string seq = find_conseq(tv);
          if (dcnt <= 2 && tv.cnt / dcnt > 5) continue;  // a couple of bogus mappings must not scoop up the clips
          if (seq != " " && is_match(seq, wupseq, -1)) {
            if (sc5pp > p) V.add_cov(p, tv.cnt);
            adj_cnt(*vref, tv);
            tv.used = true;
          }
        }
        for (size_t k = 0; k < r3.scp.size(); ++k) {
          const int sc3pp = r3.scp[k];
          std::map<int, Sclip>::iterator it = R.sc3.find(sc3pp);
          if (it == R.sc3.end() || it->second.used) continue;
          Sclip& tv = it->second;
          const std::string seq = find_conseq(tv);
          if (dcnt <= 2 && tv.cnt / dcnt > 5) continue;
          if (seq != " " && is_match(seq, vc_substr1(sanpseq, sc3pp - p), 1)) {
            if (sc3pp <= p) V.add_cov(p, tv.cnt);
            Variation* lref = sc3pp <= p ? NULL : ref_var(p);
            adj_cnt3(*vref, tv, lref);
            tv.used = true;
          }
        }
        const int pe = tail;
        Variation* h = ref_var(p);
        // taking the size of the gap into account
        if (bams && batch && pe - p >= 5 && pe - p < R.max_read_len - 10 && h != NULL && h->cnt != 0 &&
            vref->cnt > 2 * h->cnt * (1 - (pe - p) / (double)R.max_read_len) && no_passing_reads(p, pe))
          adj_cnt3(*vref, *h, h);
      } catch (const RealignAbort&) {
      }
    }
    for (int i = (int)tmp.size() - 1; i > 0; i--) {
      const int p = tmp[(size_t)i].position;
      const std::string vn = tmp[(size_t)i].desc;
      if (!V.count_pos(p)) continue;
      Variation* vref = V.find(p, vn);
      if (!vref) continue;
      // (-\d+)&[ATGC]+$
      size_t amp = vn.rfind('&');
      if (amp == std::string::npos || amp + 1 >= vn.size()) continue;
      bool tail_ok = true;
      for (size_t j = amp + 1; j < vn.size(); ++j) tail_ok = tail_ok && is_atgc(vn[j]);
      if (!tail_ok) continue;
      size_t d0 = amp;
      while (d0 > 0 && is_digit(vn[d0 - 1])) --d0;
      if (d0 == amp || d0 == 0 || vn[d0 - 1] != '-') continue;
      const std::string tn = vn.substr(d0 - 1, amp - (d0 - 1));
      Variation* tref = V.find(p, tn);
      if (tref && vref->cnt < tref->cnt) {
        adj_cnt(*tref, *vref);
        erase_ni(p, vn, false);
      }
    }
  }

  // realignins, VariationRealigner.cpp:785-1114
  std::string realign_ins(const std::map<int, std::map<std::string, int> >& pins) {
    std::vector<PosDesc> tmp = fill_and_sort(pins);
    std::string NEWINS;
    for (size_t ti = 0; ti < tmp.size(); ++ti) {
      try {
        const int position = tmp[ti].position;
        const std::string vn = tmp[ti].desc;
        const int insertion_count = tmp[ti].count;
        std::string insert1;
        if (vn.size() > 1 && vn[0] == '+' && is_atgc(vn[1])) {  // ^\+([ATGC]+).*
          size_t z = 1;
          while (z < vn.size() && is_atgc(vn[z])) ++z;
          insert1 = vn.substr(1, z - 1);
        } else {
          continue;
        }
        std::string ins3;
        int inslen = (int)insert1.size();
        if (vn.find("<dup") != std::string::npos) throw RealignAbort();  // <dup(\d+)>([ATGC]+)$ — SV keys only
        std::string extra;
        amp_atgc_local(vn, &extra);  // .*&([ATGC]+).*
        std::string compm;           // #([ATGC]+).*
        for (size_t j = 0; j + 1 < vn.size(); ++j)
          if (vn[j] == '#' && is_atgc(vn[j + 1])) {
            size_t z = j + 1;
            while (z < vn.size() && is_atgc(vn[z])) ++z;
            compm = vn.substr(j + 1, z - j - 1);
            break;
          }
        std::string newins;  // \^([ATGC]+)$
        for (size_t j = 0; j + 1 < vn.size(); ++j)
          if (vn[j] == '^') {
            bool ok = true;
            for (size_t z = j + 1; z < vn.size(); ++z) ok = ok && is_atgc(vn[z]);
            if (ok) { newins = vn.substr(j + 1); break; }
          }
        int newdel = 0;  // \^(\d+)$
        {
          size_t e = vn.size(), k = e;
          while (k > 0 && is_digit(vn[k - 1])) --k;
          if (k < e && k > 0 && vn[k - 1] == '^') newdel = atoi(vn.c_str() + k);
        }
        std::string tn = vn;
        if (!tn.empty() && tn[0] == '+') tn.erase(0, 1);
        replace_first_char(tn, '&');
        replace_first_char(tn, '#');
        {  // \^\d+$
          size_t e = tn.size(), k = e;
          while (k > 0 && is_digit(tn[k - 1])) --k;
          if (k < e && k > 0 && tn[k - 1] == '^') tn.erase(k - 1);
        }
        replace_first_char(tn, '^');
        const int wustart = position - 150 > 1 ? (position - 150) : 1;
        const std::string wupseq = join_ref_at(ref, wustart, position) + tn;
        int sanend = position + (int)vn.size() + 100;
        if (chr_len > 0 && chr_len < sanend) sanend = chr_len;
        std::string sanpseq;
        MismatchResult findmm3;
        // (ins3 is only set for <dup..> keys)
        sanpseq = tn + join_ref_at(ref, position + (int)extra.size() + 1 + (int)compm.size() + newdel, sanend);
        findmm3 = find_mm3(position + 1, sanpseq);
        MismatchResult findmm5 = find_mm5(position + (int)extra.size() + (int)compm.size() + newdel, wupseq);
        const int nm3 = findmm3.nm, nm5 = findmm5.nm;
        std::vector<Mismatch> mmm(findmm3.mm);
        mmm.insert(mmm.end(), findmm5.mm.begin(), findmm5.mm.end());
        Variation* vref = &R.ins[position][vn];  // getVariation(insertionVariants, position, vn)
        for (size_t mi = 0; mi < mmm.size(); ++mi) {
          std::string mb = mmm[mi].seq;
          const int mp = mmm[mi].pos, me = mmm[mi].end;
          if (mb.size() > 1) mb = std::string(1, mb[0]) + "&" + mb.substr(1);
          if (!V.count_pos(mp)) continue;
          Variation* variation = V.find(mp, mb);
          if (!variation) continue;
          if (variation->cnt == 0) continue;
          if (variation->sum_q / variation->cnt < P.goodq) continue;
          if (variation->sum_tp / variation->cnt > (me == 3 ? nm3 + 4 : nm5 + 4)) continue;
          if (variation->cnt >= insertion_count + (int)insert1.size() || variation->cnt / insertion_count >= 8) continue;
          if (mp > position && me == 5) V.add_cov(position, variation->cnt);
          Variation* lref = NULL;
          if (mp > position && me == 3 && V.count_pos(position) && ref.has(position))
            lref = V.find(position, std::string(1, ref_char(position)));
          if (lref) adj_cnt3(*vref, *variation, lref);
          else adj_cnt(*vref, *variation);
          erase_ni(mp, mb, true);
        }
        if (findmm3.misp != 0 && findmm3.mm.size() == 1 && V.count_pos(findmm3.misp)) {
          Variation* t = V.find(findmm3.misp, findmm3.misnt);
          if (t && t->cnt < insertion_count) erase_ni(findmm3.misp, findmm3.misnt, false);
        }
        if (findmm5.misp != 0 && findmm5.mm.size() == 1 && V.count_pos(findmm5.misp)) {
          Variation* t = V.find(findmm5.misp, findmm5.misnt);
          if (t && t->cnt < insertion_count) erase_ni(findmm5.misp, findmm5.misnt, false);
        }
        for (size_t k = 0; k < findmm5.scp.size(); ++k) {
          const int sc5pp = findmm5.scp[k];
          std::map<int, Sclip>::iterator it = R.sc5.find(sc5pp);
          if (it == R.sc5.end()) continue;
          Sclip& tv = it->second;
          if (tv.used) continue;
          const std::string seq = find_conseq(tv);
          if (seq != " " && is_match(seq, wupseq, -1)) {
            if (sc5pp > position) V.add_cov(position, tv.cnt);
            adj_cnt(*vref, tv);
            tv.used = true;
          }
        }
        for (size_t k = 0; k < findmm3.scp.size(); ++k) {
          const int sc3pp = findmm3.scp[k];
          std::map<int, Sclip>::iterator it = R.sc3.find(sc3pp);
          if (it == R.sc3.end()) continue;
          Sclip& tv = it->second;
          if (tv.used) continue;
          const std::string seq = find_conseq(tv);
          const std::string mseq = !ins3.empty() ? sanpseq : vc_substr1(sanpseq, sc3pp - position - 1);
          if (seq != " " && is_match(seq, mseq, 1)) {
            if (sc3pp <= position || (double)insert1.size() > tv.sum_tp / tv.cnt) V.add_cov(position, tv.cnt);
            Variation* lref = NULL;
            if (sc3pp > position && V.count_pos(position) && ref.has(position))
              lref = V.find(position, std::string(1, ref_char(position)));
            if ((double)insert1.size() > tv.sum_tp / tv.cnt) lref = NULL;
            adj_cnt3(*vref, tv, lref);
            tv.used = true;
            if (insert1.size() + 1 == vn.size() && (int)insert1.size() > R.max_read_len &&
                sc3pp >= position + 1 + (int)insert1.size()) {
              int flag = 0;
              const int offset = (sc3pp - position - 1) % (int)insert1.size();
              std::string tvn = vn;
              for (int seqi = 0; seqi < (int)seq.size() && seqi + offset < (int)insert1.size(); seqi++) {
                if (vc_substr2(seq, seqi, 1) != vc_substr2(insert1, seqi + offset, 1)) {
                  flag++;
                  const int shift = seqi + offset + 1;
                  tvn = tvn.substr(0, (size_t)shift) + vc_substr2(seq, seqi, 1) + tvn.substr((size_t)shift + 1);
                }
              }
              if (flag > 0) {
                Variation moved = R.ins[position][vn];
                R.ins[position][tvn] = moved;
                R.ins[position].erase(vn);
                vref = &R.ins[position][tvn];
                NEWINS = tvn;
              }
            }
          }
        }
        const int first3 = findmm3.scp[0], first5 = findmm5.scp[0];
        if (!findmm3.scp.empty() && !findmm5.scp.empty() && first3 > first5 + 3 && first3 - first5 < R.max_read_len * 0.75) {
          if (ref.has(position) && V.count_pos(position)) {
            Variation* rv = V.find(position, std::string(1, ref_char(position)));
            if (rv) adj_ref_factor(rv, (first3 - first5 - 1) / (double)R.max_read_len);
          }
          adj_ref_factor(vref, -(first3 - first5 - 1) / (double)R.max_read_len);
        }
      } catch (const RealignAbort&) {
      }
    }
    for (int i = (int)tmp.size() - 1; i > 0; i--) {
      const int p = tmp[(size_t)i].position;
      const std::string vn = tmp[(size_t)i].desc;
      std::map<int, KeyMap>::iterator pit = R.ins.find(p);
      if (pit == R.ins.end()) continue;
      KeyMap::iterator kit = pit->second.find(vn);
      if (kit == pit->second.end()) continue;
      Variation* vref = &kit->second;
      // (\+[ATGC]+)&[ATGC]+$
      size_t amp = vn.rfind('&');
      if (amp == std::string::npos || amp + 1 >= vn.size()) continue;
      bool ok = true;
      for (size_t j = amp + 1; j < vn.size(); ++j) ok = ok && is_atgc(vn[j]);
      if (!ok) continue;
      size_t b0 = amp;
      while (b0 > 0 && is_atgc(vn[b0 - 1])) --b0;
      if (b0 == amp || b0 == 0 || vn[b0 - 1] != '+') continue;
      const std::string tn = vn.substr(b0 - 1, amp - (b0 - 1));
      KeyMap::iterator tit = pit->second.find(tn);
      if (tit != pit->second.end()) {
        Variation* tref = &tit->second;
        if (vref->cnt < tref->cnt) {
          adj_cnt3(*tref, *vref, ref_var(p));
          pit->second.erase(vn);
        }
      }
    }
    return NEWINS;
  }
  static bool amp_atgc_local(const std::string& s, std::string* g1) {  // .*&([ATGC]+).* : the last '&' followed by ATGC
    for (size_t i = s.size(); i-- > 0;) {
      if (s[i] == '&' && i + 1 < s.size() && is_atgc(s[i + 1])) {
        size_t j = i + 1;
        while (j < s.size() && is_atgc(s[j])) ++j;
        *g1 = s.substr(i + 1, j - i - 1);
        return true;
      }
    }
    return false;
  }

  // find35match, VariationRealigner.cpp:1396-1430
  static void find35match(const std::string& seq5, const std::string& seq3, int* b5o, int* b3o, int* maxo) {
    const int longMismatch = 2;
    *b5o = 0; *b3o = 0; *maxo = 0;
    const int l5 = (int)seq5.size(), l3 = (int)seq3.size();
    for (int i = 0; i < l5 - 8; i++) {
      for (int j = 1; j < l3 - 8; j++) {
        int nmm = 0, total = 0;
        while (total + j <= l3 && i + total <= l5) {
          const char c3 = seq3[(size_t)(l3 - j - total)];  // vc_substr(seq3, -j - total, 1)
          const bool have5 = i + total < l5;
          if (!have5 || c3 != seq5[(size_t)(i + total)]) nmm++;
          if (nmm > longMismatch) break;
          total++;
        }
        if (total - nmm > *maxo && total - nmm > 8 && nmm / (double)total < 0.1 && (total + j >= l3 || i + total >= l5)) {
          *maxo = total - nmm;
          *b3o = j;
          *b5o = i;
          return;
        }
      }
    }
  }

  // realignlgins30, VariationRealigner.cpp:1119-1354
  void realign_lgins30() {
    struct SortSc { int position; Sclip* sc; int count; };
    struct Cmp3 { bool operator()(const SortSc& a, const SortSc& b) const {
      if (a.count != b.count) return a.count > b.count;
      return a.position < b.position; } };
    const int EXT = 5000;  // CONF_EXTENSION
    std::vector<SortSc> tmp5, tmp3;
    for (std::map<int, Sclip>::iterator it = R.sc5.begin(); it != R.sc5.end(); ++it) {
      if (it->first < R.start - EXT || it->first > R.end + EXT) continue;
      SortSc s; s.position = it->first; s.sc = &it->second; s.count = it->second.cnt;
      tmp5.push_back(s);
    }
    std::sort(tmp5.begin(), tmp5.end(), Cmp3());
    for (std::map<int, Sclip>::iterator it = R.sc3.begin(); it != R.sc3.end(); ++it) {
      if (it->first < R.start - EXT || it->first > R.end + EXT) continue;
      SortSc s; s.position = it->first; s.sc = &it->second; s.count = it->second.cnt;
      tmp3.push_back(s);
    }
    std::sort(tmp3.begin(), tmp3.end(), Cmp3());
    const int maxrl = R.max_read_len;
    for (size_t a = 0; a < tmp5.size(); ++a) {
      const int p5 = tmp5[a].position;
      Sclip* sc5v = tmp5[a].sc;
      const int cnt5 = tmp5[a].count;
      if (sc5v->used) continue;
      const std::string seq5 = find_conseq(*sc5v);
      if (seq5.size() <= 10) continue;
      for (size_t b = 0; b < tmp3.size(); ++b) {
        try {
          const int p3 = tmp3[b].position;
          Sclip* sc3v = tmp3[b].sc;
          const int cnt3 = tmp3[b].count;
          if (sc5v->used) break;
          if (sc3v->used) continue;
          if (p5 - p3 > maxrl * 2.5) continue;
          if (p3 - p5 > maxrl - 10) continue;  // if they're too far away, don't even try
          const std::string seq3 = find_conseq(*sc3v);
          if (seq3.size() <= 10) continue;
          if (!(cnt5 / (double)cnt3 >= 0.08 && cnt5 / (double)cnt3 <= 12)) continue;
          int bp5, bp3, score;
          find35match(seq5, seq3, &bp5, &bp3, &score);
          if (score == 0) continue;
          const int smscore = score / 2;  // higher quality bases: read ends are usually poor
          std::string ins = bp3 + smscore > 1 ? vc_substr2(seq3, 0, -(bp3 + smscore) + 1) : seq3;
          if (bp5 + smscore > 0) {
            std::string t = seq5.substr(0, (size_t)(bp5 + smscore));
            std::reverse(t.begin(), t.end());
            ins += t;
          }
          if (is_low_complex(ins)) continue;
          int bi = 0;
          Variation* vref;
          const int l3 = (int)seq3.size(), l5 = (int)seq5.size();
          if (seq3.size() > ins.size() &&
              !is_match(seq3.substr(ins.size()), join_ref_at(ref, p5, p5 + l3 - (int)ins.size() + 2), 1))
            continue;
          if (seq5.size() > ins.size() &&
              !is_match(seq5.substr(ins.size()), join_ref_at(ref, p3 - (l5 - (int)ins.size()) - 2, p3 - 1), -1))
            continue;
          if (p5 > p3) {
            const std::string tmp = join_ref_at(ref, p3, p5 - 1);
            if (tmp.size() > ins.size()) {  // deletion is longer
              ins = std::to_string(p3 - p5) + "^" + ins;
              bi = p3;
              vref = &V.keys(p3)[ins];
            } else if (tmp.size() < ins.size()) {
              ins = vc_substr2(ins, 0, (int)ins.size() - (int)tmp.size()) + "&" + vc_substr1(ins, p3 - p5);
              ins = "+" + ins;
              bi = p3 - 1;
              vref = &R.ins[bi][ins];
            } else {  // long MNP
              ins = "-" + std::to_string(ins.size()) + "^" + ins;
              bi = p3;
              vref = &V.keys(p3)[ins];
            }
          } else {
            std::string tmp;
            if ((int)ins.size() <= p3 - p5) {  // tandem duplication
              int rpt = 2, tnr = 3;
              const size_t span = (size_t)(p3 - p5) + ins.size();
              while ((span / (double)tnr) / (double)ins.size() > 1) {
                if (span % (size_t)tnr == 0) rpt++;
                tnr++;
              }
              tmp += join_ref_double(ref, p5, (p5 + span / (double)rpt - ins.size()));
              ins = "+" + tmp + ins;
            } else {
              tmp += join_ref_at(ref, p5, p3 - 1);
              if ((ins.size() - tmp.size()) % 2 == 0) {
                const int tex = (int)((ins.size() - tmp.size()) / 2);
                ins = (tmp + vc_substr2(ins, 0, tex)) == vc_substr1(ins, tex) ? ("+" + vc_substr1(ins, tex)) : "+" + tmp + ins;
              } else {
                ins = "+" + tmp + ins;
              }
            }
            bi = p5 - 1;
            vref = &R.ins[bi][ins];
          }
          sc3v->used = true;
          sc5v->used = true;
          vref->pstd = true;
          vref->qstd = true;
          V.add_cov(bi, sc5v->cnt);
          if (ins[0] == '+') {
            Variation* mvref = ref_var(bi);
            adj_cnt3(*vref, *sc3v, mvref);
            adj_cnt(*vref, *sc5v);
            if (batch && p3 - p5 >= 5 && p3 - p5 < maxrl - 10 && mvref != NULL && mvref->cnt != 0 &&
                vref->cnt > 2 * mvref->cnt && no_passing_reads(p5, p3))
              adj_cnt3(*vref, *mvref, mvref);
            std::map<int, std::map<std::string, int> > tins;
            tins[bi][ins] = vref->cnt;
            realign_ins(tins);
          } else if (ins[0] == '-') {
            adj_cnt3(*vref, *sc3v, ref_var(bi));
            adj_cnt(*vref, *sc5v);
            std::map<int, std::map<std::string, int> > tdel;
            tdel[bi][ins] = vref->cnt;
            realign_del(false, tdel);
          } else {
            adj_cnt(*vref, *sc3v);
            adj_cnt(*vref, *sc5v);
          }
          break;
        } catch (const RealignAbort&) {
        }
      }
    }
  }

  void realign_indels() {  // VariationRealigner.cpp:467-484
    realign_del(true, std::map<int, std::map<std::string, int> >(R.pdel));
    realign_ins(R.pins);
    realign_lgins30();
  }
};

struct RealignState {
  std::set<int> cov_touched;
  std::set<std::pair<int, char> > erased_dense;
};

inline void realign_region(const rv_params& P, RegionPileup& R, const rvk::RefView& ref, int chr_len,
                           RealignState* st = NULL, const ReadBatch* batch = NULL, int64_t read_lo = 0, int64_t read_hi = 0) {
  NiView V(R);
  adjust_mnp(V, ref);
  if (P.local_realign) {  // VariationRealigner.cpp:147-151
    Realigner rl(P, R, V, ref, chr_len, batch, read_lo, read_hi);
    rl.realign_indels();
  }
  if (st) { st->cov_touched = V.cov_touched; st->erased_dense = V.erased_dense; }
  R.cov_touched.insert(V.cov_touched.begin(), V.cov_touched.end());
  R.erased_dense.insert(V.erased_dense.begin(), V.erased_dense.end());
}

// keys longer than RV_PATCH_KEY_MAX cannot travel to the scoring kernels: dropped and counted (reported by the CLI)
inline std::atomic<long long>& dropped_patch_keys() {
  static std::atomic<long long> n(0);
  return n;
}

inline bool fill_patch(rv_patch_entry& e, int region, int pos, int table, const std::string& key, const Variation& v) {
  if (key.size() > sizeof e.key) {
    dropped_patch_keys()++;
    return false;
  }
  memset(&e, 0, sizeof e);
  e.region = region;
  e.pos = pos;
  e.table = (uint8_t)table;
  e.keylen = (uint8_t)std::min<size_t>(key.size(), sizeof e.key);
  memcpy(e.key, key.data(), e.keylen);
  e.v.cnt = v.cnt; e.v.fwd = v.fwd; e.v.rev = v.rev; e.v.lo = v.lo; e.v.hi = v.hi; e.v.extracnt = v.extracnt;
  e.v.sum_tp = v.sum_tp; e.v.sum_q = v.sum_q; e.v.sum_mapq = v.sum_mapq; e.v.sum_nm = v.sum_nm;
  e.v.pstd = v.pstd; e.v.qstd = v.qstd;
  return true;
}

// All sparse keys of the region grouped by position: non-insertion keys (table 0), tombstones of erased
// dense keys (table 2), then insertion keys (table 1); each table in key order.
inline void build_patch(const RegionPileup& R, std::vector<rv_patch_entry>* out) {
  out->clear();
  std::set<int> positions;
  for (std::map<int, KeyMap>::const_iterator a = R.ni.begin(); a != R.ni.end(); ++a) positions.insert(a->first);
  for (std::map<int, KeyMap>::const_iterator a = R.ins.begin(); a != R.ins.end(); ++a) positions.insert(a->first);
  for (std::set<std::pair<int, char> >::const_iterator e = R.erased_dense.begin(); e != R.erased_dense.end(); ++e)
    positions.insert(e->first);
  for (std::set<int>::const_iterator p = positions.begin(); p != positions.end(); ++p) {
    const int pos = *p;
    std::map<int, KeyMap>::const_iterator a = R.ni.find(pos);
    if (a != R.ni.end())
      for (KeyMap::const_iterator k = a->second.begin(); k != a->second.end(); ++k) {
        rv_patch_entry e;
        if (fill_patch(e, R.region_idx, pos, 0, k->first, k->second)) out->push_back(e);
      }
    for (int al = 0; al < 4; ++al) {
      char b = "ACGT"[al];
      if (R.erased_dense.count(std::make_pair(pos, b)) && !(a != R.ni.end() && a->second.count(std::string(1, b)))) {
        rv_patch_entry e;
        fill_patch(e, R.region_idx, pos, 2, std::string(1, b), Variation());
        out->push_back(e);
      }
    }
    std::map<int, KeyMap>::const_iterator b = R.ins.find(pos);
    if (b != R.ins.end())
      for (KeyMap::const_iterator k = b->second.begin(); k != b->second.end(); ++k) {
        rv_patch_entry e;
        if (fill_patch(e, R.region_idx, pos, 1, k->first, k->second)) out->push_back(e);
      }
  }
}

inline void collect_cov_patch(RegionPileup& R, std::vector<int32_t>* reg, std::vector<int32_t>* pos,
                              std::vector<int32_t>* val) {
  for (std::set<int>::const_iterator p = R.cov_touched.begin(); p != R.cov_touched.end(); ++p) {
    if (!R.in_table(*p)) continue;
    reg->push_back(R.region_idx);
    pos->push_back(*p);
    val->push_back((int32_t)R.cov_at(*p));
  }
}

}  // namespace rvhost
