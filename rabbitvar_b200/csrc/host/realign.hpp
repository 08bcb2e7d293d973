// realign.hpp — host-side stage between pileup and scoring (VariationRealigner::process,
// reference src/VariationRealigner.cpp:135-163) operating on the RegionPileup model, and the
// write-back of its result as patch entries for rv_apply_patch.
#pragma once
#include "pileup_model.hpp"
#include "../kernels/rv_core.cuh"
#include <string.h>

namespace rvhost {

inline void realign_region(const rv_params& P, RegionPileup& R, const rvk::RefView& ref, int chr_len) {
  (void)P; (void)R; (void)ref; (void)chr_len;
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

// All sparse keys of the region, grouped by position (non-insertion keys first, each table in key order).
inline void build_patch(const RegionPileup& R, std::vector<rv_patch_entry>* out) {
  out->clear();
  std::map<int, KeyMap>::const_iterator a = R.ni.begin(), b = R.ins.begin();
  while (a != R.ni.end() || b != R.ins.end()) {
    int pos;
    if (b == R.ins.end() || (a != R.ni.end() && a->first <= b->first)) pos = a->first;
    else pos = b->first;
    if (a != R.ni.end() && a->first == pos) {
      for (KeyMap::const_iterator k = a->second.begin(); k != a->second.end(); ++k) {
        rv_patch_entry e;
        fill_patch(e, R.region_idx, pos, 0, k->first, k->second);
        out->push_back(e);
      }
      ++a;
    }
    if (b != R.ins.end() && b->first == pos) {
      for (KeyMap::const_iterator k = b->second.begin(); k != b->second.end(); ++k) {
        rv_patch_entry e;
        fill_patch(e, R.region_idx, pos, 1, k->first, k->second);
        out->push_back(e);
      }
      ++b;
    }
  }
}

inline void collect_cov_patch(const RegionPileup& R, std::vector<int32_t>* reg, std::vector<int32_t>* pos,
                              std::vector<int32_t>* val) {
  (void)R; (void)reg; (void)pos; (void)val;
}

}  // namespace rvhost
