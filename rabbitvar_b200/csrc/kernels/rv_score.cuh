// rv_score.cuh — per-position scoring (ToVarsBuilder::process, src/ToVarsBuilder.cpp:97-205) and the
// Fisher exact test (htslib kfunc.c kt_fisher_exact; call sites simpleMode.cpp:98, somaticMode.cpp:132)
// as __host__ __device__ code; launched one position per thread from rv_abi.cu.
#pragma once
#include "rv_core.cuh"
#include <math.h>

namespace rvk {

// strandBias, include/VariationUtils.h:511-521
RV_HD int strand_bias(int fwd, int rev, const rv_params& P) {
  if (fwd + rev <= 12) return fwd * rev > 0 ? 2 : 0;
  return (fwd / (double)(fwd + rev) >= P.bias && rev / (double)(fwd + rev) >= P.bias && fwd >= P.min_bias_reads &&
          rev >= P.min_bias_reads) ? 2 : 1;
}

// A sequence given piecewise: (optional) literal key chars followed/preceded by reference spans.
// findMSI (ToVarsBuilder.cpp:680-722) only ever indexes its three strings, so they are modelled as
// concatenations of up to two spans each: a reference span [a, b] and a literal span.
struct Span {
  const char* lit;  // literal chars (key), may be null
  int lit_n;
  int ref_a, ref_n;  // reference span start position and length (after the literal if lit_first)
  bool lit_first;
  RV_HD int len() const { return lit_n + ref_n; }
};
template <class RefT>
RV_HD char span_at(const Span& s, const RefT& ref, int i) {
  if (s.lit_first) {
    if (i < s.lit_n) return s.lit[i];
    return ref.at(s.ref_a + (i - s.lit_n));
  }
  if (i < s.ref_n) return ref.at(s.ref_a + i);
  return s.lit[i - s.ref_n];
}
RV_HD Span ref_span(int a, int b) {  // joinRef(ref, a, b): positions a..b inclusive (empty when b < a)
  Span s;
  s.lit = 0;
  s.lit_n = 0;
  s.ref_a = a;
  s.ref_n = b >= a ? b - a + 1 : 0;
  s.lit_first = false;
  return s;
}
RV_HD Span lit_span(const char* p, int n) {
  Span s;
  s.lit = p;
  s.lit_n = n;
  s.ref_a = 0;
  s.ref_n = 0;
  s.lit_first = true;
  return s;
}

// Reference view without bounds tests, for callers that proved the whole span lies inside the loaded window.
struct RefRaw {
  const char* b;  // b[p] = base at reference position p
  RV_HD char at(int p) const { return b[p]; }
};

struct Msi {
  double msi;
  int shift3;
  int msint_len;
};

// findMSI(tseq1, tseq2, left), ToVarsBuilder.cpp:656-722 — the off-by-one guards are reproduced as written.
template <class RefT>
RV_HDN Msi find_msi(const Span& tseq1, const Span& tseq2, const Span& left, const RefT& ref) {
  int nmsi = 1, shift3 = 0, best_len = 0;
  double msicnt = 0;
  const int l1 = tseq1.len(), l2 = tseq2.len(), ll = left.len();
  const int lm = ll + l1;  // "mubiao" = left + tseq1
  while (nmsi <= l1 && nmsi <= 6) {
    // msint = last nmsi chars of tseq1
    // curmsi_from_end(mubiao, msint, nmsi) :656-666
    double curmsi = 1.0;
    for (int i = lm - nmsi; i > 0; i -= nmsi) {
      bool eq = i > nmsi;
      if (eq) {
        for (int k = 0; k < nmsi && eq; ++k) {
          int mi = i - nmsi + k;  // index in left+tseq1
          char a = mi < ll ? span_at(left, ref, mi) : span_at(tseq1, ref, mi - ll);
          char b = span_at(tseq1, ref, l1 - nmsi + k);
          eq = a == b;
        }
      }
      if (eq) curmsi += 1;
      else break;
    }
    for (int i = 0; i < l2; i += nmsi) {
      bool eq = i + nmsi < l2;
      if (eq)
        for (int k = 0; k < nmsi && eq; ++k) eq = span_at(tseq2, ref, i + k) == span_at(tseq1, ref, l1 - nmsi + k);
      if (eq) curmsi += 1.0;
      else break;
    }
    if (curmsi > msicnt) {
      best_len = nmsi;
      msicnt = curmsi;
    }
    nmsi++;
  }
  // shift3: common prefix of tseq1+tseq2 and tseq2
  while (shift3 < l2) {
    char a = shift3 < l1 ? span_at(tseq1, ref, shift3) : span_at(tseq2, ref, shift3 - l1);
    if (a != span_at(tseq2, ref, shift3)) break;
    shift3++;
  }
  Msi m;
  m.msi = msicnt;
  m.shift3 = shift3;
  m.msint_len = best_len;
  return m;
}

// ---- Fisher exact test -----------------------------------------------------------------------------
// lgt[n] = lgamma(n+1) table, valid for n < lgt_n; falls back to lgamma() above it.
struct LgTable {
  const double* t;
  int n;
  RV_HD double lg1(int k) const { return (t && k < n) ? t[k] : lgamma((double)k + 1.0); }
};
RV_HD double lbinom(const LgTable& L, int n, int k) {
  if (k == 0 || n == k) return 0;
  return L.lg1(n) - L.lg1(k) - L.lg1(n - k);
}
RV_HD double hypergeo(const LgTable& L, int n11, int n1_, int n_1, int n) {
  return exp(lbinom(L, n1_, n11) + lbinom(L, n - n1_, n_1 - n11) - lbinom(L, n, n_1));
}
struct HgAcc { int n11, n1_, n_1, n; double p; };
RV_HD double hypergeo_acc(const LgTable& L, int n11, int n1_, int n_1, int n, HgAcc* aux) {
  if (n1_ || n_1 || n) {
    aux->n11 = n11; aux->n1_ = n1_; aux->n_1 = n_1; aux->n = n;
  } else {
    if (n11 % 11 && n11 + aux->n - aux->n1_ - aux->n_1) {
      if (n11 == aux->n11 + 1) {
        aux->p *= (double)(aux->n1_ - aux->n11) / n11 * (aux->n_1 - aux->n11) / (n11 + aux->n - aux->n1_ - aux->n_1);
        aux->n11 = n11;
        return aux->p;
      }
      if (n11 == aux->n11 - 1) {
        aux->p *= (double)aux->n11 / (aux->n1_ - n11) * (aux->n11 + aux->n - aux->n1_ - aux->n_1) / (aux->n_1 - n11);
        aux->n11 = n11;
        return aux->p;
      }
    }
    aux->n11 = n11;
  }
  aux->p = hypergeo(L, aux->n11, aux->n1_, aux->n_1, aux->n);
  return aux->p;
}
RV_HDN void fisher_exact(const LgTable& L, int n11, int n12, int n21, int n22, double* left_o, double* right_o,
                         double* two_o) {
  int i, j, max, min;
  double p, q, left, right;
  HgAcc aux;
  int n1_ = n11 + n12, n_1 = n11 + n21, n = n11 + n12 + n21 + n22;
  max = (n_1 < n1_) ? n_1 : n1_;
  min = n1_ + n_1 - n;
  if (min < 0) min = 0;
  *two_o = *left_o = *right_o = 1.;
  if (min == max) return;
  q = hypergeo_acc(L, n11, n1_, n_1, n, &aux);
  p = hypergeo_acc(L, min, 0, 0, 0, &aux);
  for (left = 0., i = min + 1; p < 0.99999999 * q && i <= max; ++i) left += p, p = hypergeo_acc(L, i, 0, 0, 0, &aux);
  --i;
  if (p < 1.00000001 * q) left += p;
  else --i;
  p = hypergeo_acc(L, max, 0, 0, 0, &aux);
  for (right = 0., j = max - 1; p < 0.99999999 * q && j >= 0; --j) right += p, p = hypergeo_acc(L, j, 0, 0, 0, &aux);
  ++j;
  if (p < 1.00000001 * q) right += p;
  else ++j;
  *two_o = left + right;
  if (*two_o > 1.) *two_o = 1.;
  int di = i - n11, dj = j - n11;
  if ((di < 0 ? -di : di) < (dj < 0 ? -dj : dj)) right = 1. - left + q;
  else left = 1.0 - right + q;
  *left_o = left;
  *right_o = right;
}

// ---- per-position scoring ----------------------------------------------------------------------------
static const int RV_MAX_KEYS = 24;

struct KeyAcc {   // one (key, Variation) at the position, reference field set
  int cnt, fwd, rev, lo, hi, extracnt;
  double sum_tp, sum_q, sum_mapq, sum_nm;
  bool pstd, qstd;
  bool ins;       // from insertionVariants
  int key_kind;   // 0 dense allele, 1 patch entry
  int key_id;
  const char* key;  // chars of the key (1 char for dense)
  int keylen;
};

RV_HD void acc_from_dense(KeyAcc& a, const uint32_t* row, int allele, const char* base_chars) {
  a.fwd = (int)row[RV_F_FWD];
  a.rev = (int)row[RV_F_REV];
  a.cnt = a.fwd + a.rev;
  a.hi = (int)row[RV_F_HI];
  a.lo = a.cnt - a.hi;
  a.extracnt = 0;
  a.sum_tp = (double)(int)row[RV_F_SUM_TP];
  a.sum_q = (double)(int)row[RV_F_SUM_Q];
  a.sum_mapq = (double)(int)row[RV_F_SUM_MAPQ];
  a.sum_nm = (double)(int)row[RV_F_SUM_NM];
  a.pstd = (row[RV_F_STD] >> 24) & 1;
  a.qstd = (row[RV_F_STD] >> 25) & 1;
  a.ins = false;
  a.key_kind = 0;
  a.key_id = allele;
  a.key = base_chars + allele;
  a.keylen = 1;
}
RV_HD bool dense_exists(const uint32_t* row) {
  return (row[RV_F_FWD] | row[RV_F_REV] | row[RV_F_SUM_TP] | row[RV_F_SUM_Q] | row[RV_F_SUM_MAPQ] | row[RV_F_SUM_NM] |
          row[RV_F_HI] | row[RV_F_STD]) != 0;
}
RV_HD void acc_from_patch(KeyAcc& a, const rv_patch_entry& e, int idx) {
  a.cnt = e.v.cnt; a.fwd = e.v.fwd; a.rev = e.v.rev; a.lo = e.v.lo; a.hi = e.v.hi; a.extracnt = e.v.extracnt;
  a.sum_tp = e.v.sum_tp; a.sum_q = e.v.sum_q; a.sum_mapq = e.v.sum_mapq; a.sum_nm = e.v.sum_nm;
  a.pstd = e.v.pstd != 0; a.qstd = e.v.qstd != 0;
  a.ins = e.table == 1;
  a.key_kind = 1;
  a.key_id = idx;
  a.key = e.key;
  a.keylen = e.keylen;
}
RV_HD int key_cmp(const char* a, int an, const char* b, int bn) {  // std::string::compare
  int n = an < bn ? an : bn;
  for (int i = 0; i < n; ++i) {
    unsigned char x = (unsigned char)a[i], y = (unsigned char)b[i];
    if (x != y) return x < y ? -1 : 1;
  }
  return an < bn ? -1 : (an > bn ? 1 : 0);
}

struct ScoredVar {
  KeyAcc* k;
  double freq, pmean, qual, mapq, qratio, hifreq, extrafreq, nm;
  int bias, hicnt, hicov;
};

// Emit concept:  void emit(const rv_variant&)
template <class Emit>
RV_HDN void score_position(const rv_params& P, const rv_region& R, int region_idx, int pos, const RefView& ref,
                           const uint32_t* rows /*4 alleles x 8*/, uint32_t cov_p, bool has_next,
                           const uint32_t* rows_next, uint32_t cov_next, const rv_patch_entry* patch, int patch_first,
                           int patch_n, int patch_first_next, int patch_n_next, const LgTable& lgt, Emit& out,
                           int* unsupported) {
  static const char BASES[5] = "ACGT";
  KeyAcc keys[RV_MAX_KEYS];
  int nk = 0, n_ni = 0, n_ins = 0;
  // nonInsertionVariants[pos]: dense alleles (unless shadowed by a patch entry with the same 1-char key) + patch NI
  for (int a = 0; a < 4; ++a) {
    bool shadow = false;
    for (int j = 0; j < patch_n; ++j) {
      const rv_patch_entry& e = patch[patch_first + j];
      if (e.table != 1 && e.keylen == 1 && e.key[0] == BASES[a]) shadow = true;  // 2 = tombstone
    }
    if (shadow) continue;
    if (dense_exists(rows + a * RV_ROW_U32)) {
      acc_from_dense(keys[nk], rows + a * RV_ROW_U32, a, BASES);
      nk++;
      n_ni++;
    }
  }
  for (int j = 0; j < patch_n; ++j) {
    const rv_patch_entry& e = patch[patch_first + j];
    if (e.table == 2) continue;  // key erased by the realigner
    if (nk >= RV_MAX_KEYS) { (*unsupported)++; break; }
    acc_from_patch(keys[nk], e, patch_first + j);
    if (e.table == 0) n_ni++;
    else n_ins++;
    nk++;
  }
  // ToVarsBuilder.cpp:103-136
  if (n_ni == 0) return;                       // position not in nonInsertionVariants: never visited (A-24)
  if (cov_p == 0) return;                      // :128
  const char refb = ref.has(pos) ? ref.at(pos) : (char)0;
  if (n_ni == 1 && n_ins == 0 && refb) {       // isTheSameVariationOnRef :213-233
    bool only_ref = false;
    for (int i = 0; i < nk; ++i)
      if (!keys[i].ins && keys[i].keylen == 1 && keys[i].key[0] == refb) only_ref = true;
    if (only_ref && !P.pileup && !P.has_bam2) return;
  }
  int tcov = (int)cov_p;
  int hicov = 0;  // calcHicov :570-587 (all keys, including zero-count ones)
  for (int i = 0; i < nk; ++i) hicov += keys[i].hi;

  // sort keys ascending (CMP_KEY) within each table: non-insertion first, then insertions (:159-166)
  int order[RV_MAX_KEYS];
  int no = 0;
  for (int pass = 0; pass < 2; ++pass) {
    int beg = no;
    for (int i = 0; i < nk; ++i)
      if ((keys[i].ins ? 1 : 0) == pass) {
        int j = no++;
        order[j] = i;
        while (j > beg && key_cmp(keys[order[j]].key, keys[order[j]].keylen, keys[order[j - 1]].key,
                                  keys[order[j - 1]].keylen) < 0) {
          int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t;
          --j;
        }
      }
  }
  ScoredVar var[RV_MAX_KEYS];
  int nv = 0;
  for (int oi = 0; oi < no; ++oi) {
    KeyAcc& c = keys[order[oi]];
    if (!c.ins && c.cnt == 0) continue;  // createVariant :469-472 (insertions have no such test)
    ScoredVar& v = var[nv];
    int ttcov = tcov;
    if (!c.ins) {  // createVariant :466-536
      if (c.cnt > tcov && c.extracnt > 0 && c.cnt - tcov < c.extracnt) ttcov = c.cnt;
    } else {       // createInsertion :358-452
      if (c.cnt > tcov && c.extracnt != 0 && c.cnt - tcov < c.extracnt) ttcov = c.cnt;
      if (ttcov < c.cnt) {
        ttcov = c.cnt;
        if (has_next && ref.has(pos + 1) && ttcov < (int)cov_next - c.cnt) {
          ttcov = (int)cov_next;
          (*unsupported)++;  // the reference also edits fwd/rev of pos+1's ref allele here (A-14); not reproduced
        }
        tcov = ttcov;
      }
      if (hicov < c.hi) hicov = c.hi;
    }
    v.k = &c;
    v.bias = strand_bias(c.fwd, c.rev, P);
    v.qual = c.sum_q / c.cnt;
    v.mapq = c.sum_mapq / (double)c.cnt;
    v.hicnt = c.hi;
    v.freq = c.cnt / (double)ttcov;
    v.pmean = c.sum_tp / (double)c.cnt;
    v.qratio = c.hi / (c.lo != 0 ? (double)c.lo : 0.5);
    v.hifreq = hicov > 0 ? c.hi / (double)hicov : 0;
    v.extrafreq = c.extracnt != 0 ? c.extracnt / (double)ttcov : 0;
    v.nm = c.sum_nm / (double)c.cnt;
    v.hicov = hicov;
    nv++;
  }
  // sort(var, CMP_VARI) :167 — qual*cnt descending, ties (|d| < 1e-5) by key ascending
  int vord[RV_MAX_KEYS];
  for (int i = 0; i < nv; ++i) {
    int j = i;
    vord[j] = i;
    while (j > 0) {
      ScoredVar& a = var[vord[j]];
      ScoredVar& b = var[vord[j - 1]];
      double res = a.qual * a.k->cnt - b.qual * b.k->cnt;
      bool before;
      if (res < 0.00001 && res > -0.00001) before = key_cmp(a.k->key, a.k->keylen, b.k->key, b.k->keylen) < 0;
      else before = res > 0;
      if (!before) break;
      int t = vord[j]; vord[j] = vord[j - 1]; vord[j - 1] = t;
      --j;
    }
  }
  // collectVarsAtPosition :327-346
  int ref_i = -1;
  double maxfreq = 0;
  int n_variants = 0;
  for (int oi = 0; oi < nv; ++oi) {
    ScoredVar& v = var[vord[oi]];
    if (refb && v.k->keylen == 1 && v.k->key[0] == refb && !v.k->ins) ref_i = vord[oi];
    else {
      n_variants++;
      if (v.freq > maxfreq) maxfreq = v.freq;
    }
  }
  // note: an insertion whose key is "+X" never equals the 1-char ref base, so `!ins` above is implied
  if (!P.pileup && maxfreq <= P.freq && !P.has_bam2) return;  // :170-175
  if (nv == 0) return;
  if (P.candidates_only && !P.pileup) {
    // necessary conditions of Variant::isGoodVar (include/Variant.h:205-231); a position none of whose variants
    // can pass them prints nothing in simple mode (simpleMode.cpp:176-190)
    bool any = false;
    for (int oi = 0; oi < nv; ++oi) {
      if (vord[oi] == ref_i) continue;
      const ScoredVar& v = var[vord[oi]];
      if (v.freq >= P.freq && v.hicnt >= P.minr && v.pmean >= P.read_pos_filter && v.qual >= P.goodq &&
          v.qratio >= P.qratio)
        any = true;
    }
    if (!any) return;
    if (P.candidates_only == 2) {
      // the caller only wants to know WHERE something can be printed (paired mode: the join list of
      // rv_score_positions): one record carrying region and position, no MSI / Fisher work
      rv_variant o;
      o.region = region_idx; o.pos = pos; o.cnt = o.fwd = o.rev = o.tcov = o.hicnt = o.hicov = o.ref_fwd = o.ref_rev = 0;
      o.shift3 = o.msint = 0;
      o.freq = o.pmean = o.qual = o.mapq = o.qratio = o.hifreq = o.extrafreq = o.nm = o.msi = 0;
      o.pvalue = 1.0; o.oddratio = 0.0;
      o.bias_ref = o.bias_var = o.pstd = o.qstd = o.is_ref = o.key_kind = o.pad = 0;
      o.rank = 0;
      o.key_id = 0;
      out.emit(o);
      return;
    }
  }
  // collectReferenceVariants :730-1098 — numeric part; allele strings / genotype / flanks are host work
  int rfc = 0, rrc = 0;
  if (ref_i >= 0) { rfc = var[ref_i].k->fwd; rrc = var[ref_i].k->rev; }
  if (tcov > (int)cov_p && has_next && ref.has(pos + 1)) {  // :754-760
    // the next position's reference allele as the realigner left it: a patch entry with that key shadows the
    // dense row (table 2 = erased: the key is gone and nothing is taken over)
    const char nb = ref.at(pos + 1);
    int a = allele_of(nb);
    bool shadowed = false;
    for (int j = 0; j < patch_n_next; ++j) {
      const rv_patch_entry& e = patch[patch_first_next + j];
      if (e.table == 1 || e.keylen != 1 || e.key[0] != nb) continue;
      shadowed = true;
      if (e.table == 0) { rfc = e.v.fwd; rrc = e.v.rev; }
    }
    if (!shadowed && a >= 0 && dense_exists(rows_next + a * RV_ROW_U32)) {
      rfc = (int)rows_next[a * RV_ROW_U32 + RV_F_FWD];
      rrc = (int)rows_next[a * RV_ROW_U32 + RV_F_REV];
    }
  }
  int rank = 0;
  for (int oi = 0; oi < nv; ++oi) {
    ScoredVar& v = var[vord[oi]];
    const bool is_ref = vord[oi] == ref_i;
    rv_variant o;
    o.region = region_idx;
    o.pos = pos;
    o.cnt = v.k->cnt; o.fwd = v.k->fwd; o.rev = v.k->rev;
    o.tcov = tcov;
    o.hicnt = v.hicnt; o.hicov = v.hicov;
    o.ref_fwd = rfc; o.ref_rev = rrc;
    o.shift3 = 0; o.msint = 0; o.msi = 0;
    o.freq = v.freq; o.pmean = v.pmean; o.qual = v.qual; o.mapq = v.mapq; o.qratio = v.qratio;
    o.hifreq = v.hifreq; o.extrafreq = v.extrafreq; o.nm = v.nm;
    o.bias_ref = (uint8_t)(ref_i >= 0 ? var[ref_i].bias : 0);
    o.bias_var = (uint8_t)v.bias;
    o.pstd = v.k->pstd; o.qstd = v.k->qstd;
    o.is_ref = is_ref ? 1 : 0;
    o.key_kind = (uint8_t)v.k->key_kind;
    o.key_id = v.k->key_id;
    o.pad = 0;
    o.pvalue = 1.0; o.oddratio = 0.0;
    if (is_ref) {
      o.rank = 255;
      if (n_variants == 0) {  // :1066-1090 — no variant reads detected
        o.cnt = 0; o.freq = 0; o.fwd = 0; o.rev = 0; o.bias_var = 0;
      }
    } else {
      o.rank = (uint8_t)rank++;
      // MSI context :802-878
      const char* key = v.k->key;
      const int kl = v.k->keylen;
      Msi m;
      m.msi = 0; m.shift3 = 0; m.msint_len = 0;
      if (key[0] == '+') {
        bool plain = true;
        for (int i = 0; i < kl; ++i) if (key[i] == '&' || key[i] == '#' || key[i] == '<') plain = false;
        if (plain) {  // proceedVrefIsInsertion :284-317
          Span t1 = lit_span(key + 1, kl - 1);
          Span left = ref_span(pos - 50 > 1 ? pos - 50 : 1, pos);
          Span t2 = ref_span(pos + 1, pos + 70 > R.chr_len ? R.chr_len : pos + 70);
          Msi a = find_msi(t1, t2, left, ref);
          Span none = ref_span(1, 0);
          Msi b = find_msi(left, t2, none, ref);
          m = a;
          if (m.msi < b.msi) { m.msi = b.msi; m.msint_len = b.msint_len; }
          if (m.msi <= a.shift3 / (double)(kl - 1)) m.msi = a.shift3 / (double)(kl - 1);
        }
      } else if (key[0] == '-') {
        int dellen = 0;
        for (int i = 1; i < kl && key[i] >= '0' && key[i] <= '9'; ++i) dellen = dellen * 10 + (key[i] - '0');
        if (dellen < 1000) {  // conf->SVMINLEN; proceedVrefIsDeletion :242-275
          Span left = ref_span(pos - 70 > 1 ? pos - 70 : 1, pos - 1);
          int tend = pos + dellen + 70 > R.chr_len ? R.chr_len : pos + dellen + 70;
          Span t1 = ref_span(pos, pos + dellen - 1);
          Span t2 = ref_span(pos + dellen, tend);
          Msi a = find_msi(t1, t2, left, ref);
          Span none = ref_span(1, 0);
          Msi b = find_msi(left, t2, none, ref);
          m = a;
          if (m.msi < b.msi) { m.msi = b.msi; m.msint_len = b.msint_len; }
          if (m.msi <= a.shift3 / (double)dellen) m.msi = a.shift3 / (double)dellen;
        }
      } else {  // SNV / MNV :864-878
        Span t1 = ref_span(pos - 30 > 1 ? pos - 30 : 1, pos + 1);
        Span t2 = ref_span(pos + 2, pos + 70 > R.chr_len ? R.chr_len : pos + 70);
        Span none = ref_span(1, 0);
        m = find_msi(t1, t2, none, ref);
      }
      o.msi = m.msi;
      o.shift3 = m.shift3;
      o.msint = m.msint_len;
      // negative counts are clamped before output (adjustVariantCounts :543-562)
      int a11 = rfc < 0 ? 0 : rfc, a12 = rrc < 0 ? 0 : rrc, a21 = o.fwd < 0 ? 0 : o.fwd, a22 = o.rev < 0 ? 0 : o.rev;
      if (P.fisher) {  // print_output_variant_simple :96-108
        double l, r, two;
        fisher_exact(lgt, a11, a12, a21, a22, &l, &r, &two);
        o.pvalue = two;
        double ad = (double)a11 * a22, bc = (double)a12 * a21;
        o.oddratio = (bc != 0 && ad != 0) ? (ad > bc ? ad / bc : bc / ad) : 0.0;
      }
    }
    if (o.ref_fwd < 0) o.ref_fwd = 0;
    if (o.ref_rev < 0) o.ref_rev = 0;
    if (!is_ref) {  // adjustVariantCounts (:1061) never sees the reference variant of a position that has variants
      if (o.fwd < 0) o.fwd = 0;
      if (o.rev < 0) o.rev = 0;
    }
    out.emit(o);
  }
}

// ---- dense positions ---------------------------------------------------------------------------------
// score_position restricted to positions whose only keys are the four single-base alleles of the dense table
// (no patch entry, no insertion key): same arithmetic, fixed-size state.  The MSI context and the Fisher test
// of the non-reference alleles are left to the caller (emit_variant), which runs them densely packed.
// Emit concept:  void emit(const rv_variant&)                                   complete record
//                void emit_variant(const rv_variant&, int a11, int a12, int a21, int a22)  msi/pvalue still to fill
template <class Emit>
RV_HDN void score_dense_position(const rv_params& P, int region_idx, int pos, char refb, const uint32_t* rows,
                                 uint32_t cov_p, Emit& out) {
  static const char BASES[5] = "ACGT";
  int n_exist = 0, only_al = -1, hicov = 0;
  for (int a = 0; a < 4; ++a)
    if (dense_exists(rows + a * RV_ROW_U32)) {
      n_exist++;
      only_al = a;
      hicov += (int)rows[a * RV_ROW_U32 + RV_F_HI];  // calcHicov :570-587 (zero-count keys included)
    }
  if (n_exist == 0 || cov_p == 0) return;  // ToVarsBuilder.cpp:103-128
  if (n_exist == 1 && refb && BASES[only_al] == refb && !P.pileup && !P.has_bam2) return;  // :133, :213-233
  if (P.candidates_only && !P.pileup) {
    // integer pre-screen of the candidate cut below: a position none of whose non-reference alleles has
    // hicnt >= minr cannot print anything (most sequencing-error alleles stop here, before any double arithmetic)
    bool possible = false;
    for (int a = 0; a < 4; ++a) {
      const uint32_t* row = rows + a * RV_ROW_U32;
      if (BASES[a] != refb && dense_exists(row) && row[RV_F_FWD] + row[RV_F_REV] != 0 && (int)row[RV_F_HI] >= P.minr)
        possible = true;
    }
    if (!possible) return;
  }
  const int tcov = (int)cov_p;
  int al[4], cnt[4], bias[4];
  double qual[4], freq[4], pmean[4];
  int nv = 0;
  for (int a = 0; a < 4; ++a) {  // keys in ascending order (:159-166); createVariant :466-536
    const uint32_t* row = rows + a * RV_ROW_U32;
    if (!dense_exists(row)) continue;
    const int c = (int)(row[RV_F_FWD] + row[RV_F_REV]);
    if (c == 0) continue;
    al[nv] = a;
    cnt[nv] = c;
    bias[nv] = strand_bias((int)row[RV_F_FWD], (int)row[RV_F_REV], P);
    qual[nv] = (double)(int)row[RV_F_SUM_Q] / c;
    freq[nv] = c / (double)tcov;
    pmean[nv] = (double)(int)row[RV_F_SUM_TP] / (double)c;
    nv++;
  }
  if (nv == 0) return;
  // sort(var, CMP_VARI) :167 — qual*cnt descending, ties (|d| < 1e-5) by key ascending
  int vord[4];
  for (int i = 0; i < nv; ++i) {
    int j = i;
    vord[j] = i;
    while (j > 0) {
      const int x = vord[j], y = vord[j - 1];
      const double res = qual[x] * cnt[x] - qual[y] * cnt[y];
      bool before;
      if (res < 0.00001 && res > -0.00001) before = al[x] < al[y];
      else before = res > 0;
      if (!before) break;
      vord[j] = y;
      vord[j - 1] = x;
      --j;
    }
  }
  int ref_i = -1, n_variants = 0;
  double maxfreq = 0;
  for (int oi = 0; oi < nv; ++oi) {  // collectVarsAtPosition :327-346
    const int v = vord[oi];
    if (refb && BASES[al[v]] == refb) ref_i = v;
    else {
      n_variants++;
      if (freq[v] > maxfreq) maxfreq = freq[v];
    }
  }
  if (!P.pileup && maxfreq <= P.freq && !P.has_bam2) return;  // :170-175
  if (P.candidates_only && !P.pileup) {
    bool any = false;
    for (int v = 0; v < nv; ++v) {
      if (v == ref_i) continue;
      const uint32_t* row = rows + al[v] * RV_ROW_U32;
      const int hi = (int)row[RV_F_HI], lo = cnt[v] - hi;
      const double qratio = hi / (lo != 0 ? (double)lo : 0.5);
      if (freq[v] >= P.freq && hi >= P.minr && pmean[v] >= P.read_pos_filter && qual[v] >= P.goodq && qratio >= P.qratio)
        any = true;
    }
    if (!any) return;
  }
  int rfc = 0, rrc = 0;
  if (ref_i >= 0) {
    rfc = (int)rows[al[ref_i] * RV_ROW_U32 + RV_F_FWD];
    rrc = (int)rows[al[ref_i] * RV_ROW_U32 + RV_F_REV];
  }
  int rank = 0;
  for (int oi = 0; oi < nv; ++oi) {
    const int v = vord[oi];
    const uint32_t* row = rows + al[v] * RV_ROW_U32;
    const bool is_ref = v == ref_i;
    const int c = cnt[v], hi = (int)row[RV_F_HI], lo = c - hi;
    rv_variant o;
    o.region = region_idx;
    o.pos = pos;
    o.cnt = c; o.fwd = (int)row[RV_F_FWD]; o.rev = (int)row[RV_F_REV];
    o.tcov = tcov;
    o.hicnt = hi; o.hicov = hicov;
    o.ref_fwd = rfc; o.ref_rev = rrc;
    o.shift3 = 0; o.msint = 0; o.msi = 0;
    o.freq = freq[v]; o.pmean = pmean[v]; o.qual = qual[v];
    o.mapq = (double)(int)row[RV_F_SUM_MAPQ] / (double)c;
    o.qratio = hi / (lo != 0 ? (double)lo : 0.5);
    o.hifreq = hicov > 0 ? hi / (double)hicov : 0;
    o.extrafreq = 0;
    o.nm = (double)(int)row[RV_F_SUM_NM] / (double)c;
    o.bias_ref = (uint8_t)(ref_i >= 0 ? bias[ref_i] : 0);
    o.bias_var = (uint8_t)bias[v];
    o.pstd = (row[RV_F_STD] >> 24) & 1; o.qstd = (row[RV_F_STD] >> 25) & 1;
    o.is_ref = is_ref ? 1 : 0;
    o.key_kind = 0;
    o.key_id = al[v];
    o.pad = 0;
    o.pvalue = 1.0; o.oddratio = 0.0;
    if (o.ref_fwd < 0) o.ref_fwd = 0;
    if (o.ref_rev < 0) o.ref_rev = 0;
    if (is_ref) {
      o.rank = 255;
      if (n_variants == 0) {  // :1066-1090 — no variant reads detected
        o.cnt = 0; o.freq = 0; o.fwd = 0; o.rev = 0; o.bias_var = 0;
      }
      // adjustVariantCounts (:1061) runs over the non-reference variants only: the reference variant of a
      // position that has variants keeps negative forward / reverse counts
      out.emit(o);
    } else {
      o.rank = (uint8_t)rank++;
      const int a11 = rfc < 0 ? 0 : rfc, a12 = rrc < 0 ? 0 : rrc, a21 = o.fwd < 0 ? 0 : o.fwd, a22 = o.rev < 0 ? 0 : o.rev;
      if (o.fwd < 0) o.fwd = 0;
      if (o.rev < 0) o.rev = 0;
      out.emit_variant(o, a11, a12, a21, a22);
    }
  }
}

// MSI context of an SNV/MNV key (ToVarsBuilder.cpp:864-878) and the Fisher test of print_output_variant_simple
// (simpleMode.cpp:96-108): the part of a dense record that emit_variant left open.
template <class RefT>
RV_HDN void finish_dense_variant(const rv_params& P, int pos, int chr_len, const RefT& ref, const LgTable& lgt, int a11,
                                 int a12, int a21, int a22, double* msi, int* shift3, int* msint, double* pvalue,
                                 double* oddratio) {
  Span t1 = ref_span(pos - 30 > 1 ? pos - 30 : 1, pos + 1);
  Span t2 = ref_span(pos + 2, pos + 70 > chr_len ? chr_len : pos + 70);
  Span none = ref_span(1, 0);
  Msi m = find_msi(t1, t2, none, ref);
  *msi = m.msi;
  *shift3 = m.shift3;
  *msint = m.msint_len;
  *pvalue = 1.0;
  *oddratio = 0.0;
  if (P.fisher) {
    double l, r, two;
    fisher_exact(lgt, a11, a12, a21, a22, &l, &r, &two);
    *pvalue = two;
    const double ad = (double)a11 * a22, bc = (double)a12 * a21;
    *oddratio = (bc != 0 && ad != 0) ? (ad > bc ? ad / bc : bc / ad) : 0.0;
  }
}

}  // namespace rvk
