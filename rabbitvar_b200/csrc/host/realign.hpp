// realign.hpp — host-side stage between pileup and scoring (VariationRealigner::process, reference
// src/VariationRealigner.cpp:135-163) on the RegionPileup model, and the write-back of its result as
// patch entries for rv_apply_patch.  north_star keeps this stage on the host.
//
// Implemented: adjustMNP (:334-465) with findconseq (include/VariationUtils.h:403-509) and ismatchref
// (:2004-2024).  realignIndels (:466-1354) is the next row of SURVEY.md §8(f) and is not here yet.
#pragma once
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
  for (std::map<int, std::map<char, int> >::iterator nve = sc.nt.begin(); nve != sc.nt.end(); ++nve) {
    int pis = nve->first;
    int maxCount = 0;
    double maxQuality = 0;
    char chosen = 0;
    int totalCount = 0;
    for (std::map<char, int>::iterator ent = nve->second.begin(); ent != nve->second.end(); ++ent) {
      char cb = ent->first;
      int cc = ent->second;
      totalCount += cc;
      bool hasq = sc.seq.count(pis) && sc.seq[pis].count(cb);
      if (cc > maxCount || (hasq && sc.seq[pis][cb].sum_q > maxQuality)) {
        maxCount = cc;
        chosen = cb;
        maxQuality = sc.seq[pis][cb].sum_q;
      }
    }
    if (pis == 3 && sc.nt.size() >= 6 && totalCount / (double)sc.cnt < 0.2 && totalCount <= 2) break;
    if ((totalCount - maxCount > 2 || maxCount <= totalCount - maxCount) && maxCount / (double)totalCount < 0.8) {
      if (flag) break;
      flag = true;
    }
    total += totalCount;
    match += maxCount;
    if (chosen != 0) seqq += chosen;
  }
  std::string SEQ;
  int ntSize = (int)sc.nt.size();
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

struct RealignState {
  std::set<int> cov_touched;
  std::set<std::pair<int, char> > erased_dense;
};

inline void realign_region(const rv_params& P, RegionPileup& R, const rvk::RefView& ref, int chr_len,
                           RealignState* st = NULL) {
  (void)P; (void)chr_len;
  NiView V(R);
  adjust_mnp(V, ref);
  if (st) { st->cov_touched = V.cov_touched; st->erased_dense = V.erased_dense; }
  R.cov_touched.insert(V.cov_touched.begin(), V.cov_touched.end());
  R.erased_dense.insert(V.erased_dense.begin(), V.erased_dense.end());
}

inline void fill_patch(rv_patch_entry& e, int region, int pos, int table, const std::string& key, const Variation& v) {
  memset(&e, 0, sizeof e);
  e.region = region;
  e.pos = pos;
  e.table = (uint8_t)table;
  e.keylen = (uint8_t)std::min<size_t>(key.size(), sizeof e.key);
  memcpy(e.key, key.data(), e.keylen);
  e.v.cnt = v.cnt; e.v.fwd = v.fwd; e.v.rev = v.rev; e.v.lo = v.lo; e.v.hi = v.hi; e.v.extracnt = v.extracnt;
  e.v.sum_tp = v.sum_tp; e.v.sum_q = v.sum_q; e.v.sum_mapq = v.sum_mapq; e.v.sum_nm = v.sum_nm;
  e.v.pstd = v.pstd; e.v.qstd = v.qstd;
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
        fill_patch(e, R.region_idx, pos, 0, k->first, k->second);
        out->push_back(e);
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
        fill_patch(e, R.region_idx, pos, 1, k->first, k->second);
        out->push_back(e);
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
