// rv_core.cuh — per-read logic of the pileup path: read filters, CIGAR rewrite rules, CIGAR walk.
//
// Written from scratch as table/array code (no strings, no regex, no hash maps) so it runs one read
// per GPU thread; every function cites the reference lines whose behaviour it reproduces
// (LeiHaoa/RabbitVar: src/parseCigar.cpp, src/cigarModifier.cpp, include/VariationUtils.h).
// Functions are __host__ __device__ so tests/ can also single-step them on the CPU; the product only
// ever launches them from kernels (rv_kernels.cu) — there is no CPU execution path in the library.
#pragma once
#include "../../../include/rabbitvar_b200.h"
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#if defined(__CUDACC__)
#define RV_HD __host__ __device__ __forceinline__
#define RV_HDN __host__ __device__
#else
#define RV_HD inline
#define RV_HDN inline
#endif
// Pins a loop-invariant in a register: kernel parameters live in the constant bank and the compiler otherwise re-loads
// them (LDC) inside the per-base loops, where the load sits on the critical path of every iteration.
#if defined(__CUDA_ARCH__)
#define RV_KEEP_REG(x) asm volatile("" : "+r"(x))
#else
#define RV_KEEP_REG(x) ((void)0)
#endif

// "any lane of the warp" on the device (every lane of the warp must be there), the value itself on the host
#if defined(__CUDA_ARCH__)
#define RV_WARP_ANY(x) (__syncwarp(), __any_sync(0xffffffffu, (x)))
#else
#define RV_WARP_ANY(x) (x)
#endif

namespace rvk {

enum { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_H = 5, OP_P = 6, OP_EQ = 7, OP_X = 8, OP_B = 9 };
static const int RV_MAX_OPS = 48;
static const int CONF_LOWQUAL = 10;  // include/Configuration.h:14-27

RV_HD int iceil(double x) {  // ceil() for thresholds in the int range
  int i = (int)x;
  return (double)i < x ? i + 1 : i;
}
RV_HD int c_op(uint32_t c) { return (int)(c & 0xf); }
RV_HD int c_len(uint32_t c) { return (int)(c >> 4); }
RV_HD uint32_t c_make(int len, int op) { return ((uint32_t)len << 4) | (uint32_t)op; }

// Reference window of the region (RecordPreprocessor::makeReference, recordPreprocessor.cpp:41-78):
// a position "exists" iff it lies in [lo, hi]; the bases themselves are a flat contig slice.
struct RefView {
  const char* bases;
  int32_t base_pos;  // reference position of bases[0]
  int64_t n;
  int32_t lo, hi;
  RV_HD bool has(int p) const { return p >= lo && p <= hi && p >= base_pos && (int64_t)(p - base_pos) < n; }
  // value of ref[p] as the reference's operator[] returns it: '\0' for a missing key
  RV_HD char at(int p) const { return has(p) ? bases[p - base_pos] : (char)0; }
};

RV_HD char nt16_char(int nib) {
  // seq_nt16_str "=ACMGRSVTWYHKDBN" as two 64-bit literals (a switch diverges and costs a jump table)
  const unsigned long long lo = 0x565352474d43413dull;  // "=ACMGRSV"
  const unsigned long long hi = 0x4e42444b48595754ull;  // "TWYHKDBN"
  return (char)(((nib & 8) ? hi : lo) >> (8 * (nib & 7)));
}
RV_HD int allele_of(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }
RV_HD bool is_atgc(char c) { return c == 'A' || c == 'T' || c == 'G' || c == 'C'; }

// One aligned 32-bit word of a byte array kept in registers: the per-base loop of walk_read reads its bytes (8 packed
// bases, 4 qualities, 4 reference bases per word) from it and touches memory once per word.  Device only — the host
// build (tests/tools/rv_dump --backend sim) reads the bytes themselves.
struct WordCache {
  uintptr_t tag;
  uint32_t w;
};
RV_HD int cached_byte(const uint8_t* p, WordCache& c) {
#if defined(__CUDA_ARCH__)
  const uintptr_t a = (uintptr_t)p, al = a & ~(uintptr_t)3;
  if (al != c.tag) {
    c.w = *(const uint32_t*)al;
    c.tag = al;
  }
  return (int)((c.w >> (8 * (unsigned)(a & 3))) & 0xffu);
#else
  (void)c;
  return (int)*p;
#endif
}

struct ReadView {
  const uint8_t* seq4;
  const uint8_t* qual;
  int lseq;
  RV_HD char base(int i) const {
    if (i < 0 || i >= lseq) return (char)0;
    int b = seq4[i >> 1];
    return nt16_char((i & 1) ? (b & 15) : (b >> 4));
  }
  RV_HD int q(int i) const { return (i < 0 || i >= lseq) ? 0 : (int)qual[i]; }
};

struct Cigar {
  uint32_t op[RV_MAX_OPS];
  int n;
  bool overflow;
  RV_HD void replace(int at, int count, const uint32_t* nw, int nnew) {
    int tail = n - (at + count);
    int nn = at + nnew + tail;
    if (nn > RV_MAX_OPS) { overflow = true; return; }
    if (nnew > count) for (int k = tail - 1; k >= 0; --k) op[at + nnew + k] = op[at + count + k];
    else if (nnew < count) for (int k = 0; k < tail; ++k) op[at + nnew + k] = op[at + count + k];
    for (int k = 0; k < nnew; ++k) op[at + k] = nw[k];
    n = nn;
  }
};

// Small fixed-capacity key builder (allele description strings, include/Variant.h:24-33).
struct Key {
  alignas(4) char c[RV_EVENT_KEY_MAX];
  int n;
  bool trunc;
  RV_HD void clear() { n = 0; trunc = false; }
  RV_HD void push(char ch) { if (n < RV_EVENT_KEY_MAX) c[n++] = ch; else trunc = true; }
  RV_HD void push_int(int v) {
    char tmp[12];
    int k = 0;
    if (v == 0) tmp[k++] = '0';
    bool neg = v < 0;
    unsigned u = neg ? (unsigned)(-v) : (unsigned)v;
    while (u) { tmp[k++] = (char)('0' + u % 10); u /= 10; }
    if (neg) push('-');
    while (k) push(tmp[--k]);
  }
  RV_HD void insert_at(int idx, char ch) {  // std::string::insert(idx, 1, ch); idx <= n assumed
    if (idx > n) idx = n;
    if (n >= RV_EVENT_KEY_MAX) { trunc = true; return; }
    for (int k = n; k > idx; --k) c[k] = c[k - 1];
    c[idx] = ch;
    n++;
  }
  RV_HD void erase_first(char ch) {  // replaceFirst(s, ch, "")
    for (int k = 0; k < n; ++k)
      if (c[k] == ch) { for (int j = k; j + 1 < n; ++j) c[j] = c[j + 1]; n--; return; }
  }
  RV_HD void prepend(const char* s, int len) {
    if (n + len > RV_EVENT_KEY_MAX) { trunc = true; return; }
    for (int k = n - 1; k >= 0; --k) c[k + len] = c[k];
    for (int k = 0; k < len; ++k) c[k] = s[k];
    n += len;
  }
  RV_HD bool contains(char ch) const { for (int k = 0; k < n; ++k) if (c[k] == ch) return true; return false; }
};

// ------------------------------------------------------------------------------------------------
// CigarModifier::modifyCigar (src/cigarModifier.cpp:219-410) on op arrays.
// ------------------------------------------------------------------------------------------------
struct ModCtx {
  Cigar* cg;
  int position;
  const ReadView* rd;
  const RefView* ref;
  bool unsupported;
};

// sums used by the helpers: ALIGNED_LENGTH_MND "(\d+)[MND]" and SOFT_CLIPPED "(\d+)[MIS]" over a prefix
RV_HD int sum_mnd(const Cigar& c, int upto) {
  int s = 0;
  for (int k = 0; k < upto; ++k) { int o = c_op(c.op[k]); if (o == OP_M || o == OP_N || o == OP_D) s += c_len(c.op[k]); }
  return s;
}
RV_HD int sum_mis(const Cigar& c, int upto) {
  int s = 0;
  for (int k = 0; k < upto; ++k) { int o = c_op(c.op[k]); if (o == OP_M || o == OP_I || o == OP_S) s += c_len(c.op[k]); }
  return s;
}
RV_HD bool has_eq(const RefView& r, int p, char ch) { return r.has(p) && r.at(p) == ch; }   // isHasAndEquals
RV_HD bool has_ne(const RefView& r, int p, char ch) { return r.has(p) && r.at(p) != ch; }   // isHasAndNotEquals

// distinct reference bases collected in a std::set<char> — we only need "more than one distinct"
struct BaseSet {
  unsigned mask;
  RV_HD BaseSet() : mask(0) {}
  RV_HD void add(char ch) { mask |= 1u << ((unsigned char)ch & 31); }  // A,C,G,T,N map to distinct bits
  RV_HD int size() const { int s = 0; for (unsigned m = mask; m; m &= m - 1) ++s; return s; }
};

// backward walk shared by captureMisSoftly3Mismatches / captureMisSoftlyMS (:546-562, :639-655):
// rn = 1 + index of the last mismatch seen before 3 consecutive matches, scanning from the M end.
RV_HD int tail_mismatch_walk(const ModCtx& m, int mch, int refoff, int rdoff) {
  int rn = 0, rrn = 0, rmch = 0;
  while (rrn < mch && rn < mch) {
    if (!m.ref->has(refoff - rrn - 1)) break;
    if (rrn < rdoff) {
      if (m.ref->at(refoff - rrn - 1) != m.rd->base(rdoff - rrn - 1)) { rn = rrn + 1; rmch = 0; }
      else rmch++;
    }
    rrn++;
    if (rmch >= 3) break;
  }
  return rn;
}
// forward walk shared by combineBeginDigM / combineDigSDigM (:418-433, :499-514)
RV_HD int head_mismatch_walk(const ModCtx& m, int mch, int position, int soft) {
  int rn = 0, rrn = 0, rmch = 0;
  while (rrn < mch && rn < mch) {
    if (!m.ref->has(position + rrn)) break;
    if (has_ne(*m.ref, position + rrn, m.rd->base(soft + rrn))) { rn = rrn + 1; rmch = 0; }
    else if (has_eq(*m.ref, position + rrn, m.rd->base(soft + rrn))) rmch++;
    rrn++;
    if (rmch >= 3) break;
  }
  return rn;
}

// tail "...(m)M(s)S$" : captureMisSoftlyMS, cigarModifier.cpp:574-668
RV_HD void capture_mis_softly_ms(ModCtx& m) {
  Cigar& c = *m.cg;
  int im = c.n - 2;
  int mch = c_len(c.op[im]), soft = c_len(c.op[im + 1]);
  int refoff = m.position + mch + sum_mnd(c, im);
  int rdoff = mch + sum_mis(c, im);
  int rn = 0;
  while (rn < soft && has_eq(*m.ref, refoff + rn, m.rd->base(rdoff + rn)) && m.rd->q(rdoff + rn) > CONF_LOWQUAL) rn++;
  if (rn > 0) {
    mch += rn;
    soft -= rn;
    rn = 0;
  }
  if (soft > 0) {
    BaseSet RN;
    while (rn + 1 < soft && has_eq(*m.ref, refoff + rn + 1, m.rd->base(rdoff + rn + 1)) &&
           m.rd->q(rdoff + rn + 1) > CONF_LOWQUAL) {
      rn++;
      if (m.ref->has(refoff + rn + 1)) RN.add(m.ref->at(refoff + rn + 1));
    }
    if (rn > 4 && RN.size() > 1) {
      mch += rn + 1;
      soft -= rn + 1;
    }
    if (rn == 0) {
      rn = tail_mismatch_walk(m, mch, refoff, rdoff);
      if (rn > 0 && rn < mch) {
        soft += rn;
        mch -= rn;
      }
    }
  }
  uint32_t nw[2];
  int k = 0;
  nw[k++] = c_make(mch, OP_M);
  if (soft > 0) nw[k++] = c_make(soft, OP_S);
  c.replace(im, 2, nw, k);
}

// tail "...(m)M$" : captureMisSoftly3Mismatches, cigarModifier.cpp:529-568
RV_HD void capture_mis_softly_3mm(ModCtx& m) {
  Cigar& c = *m.cg;
  int im = c.n - 1;
  int mch = c_len(c.op[im]);
  int refoff = m.position + mch + sum_mnd(c, im);
  int rdoff = mch + sum_mis(c, im);
  int rn = tail_mismatch_walk(m, mch, refoff, rdoff);
  mch -= rn;
  if (rn > 0 && rn <= 3) {
    uint32_t nw[2] = {c_make(mch, OP_M), c_make(rn, OP_S)};
    c.replace(im, 1, nw, 2);
  }
}

// head "^(s)S(m)M" : combineDigSDigM, cigarModifier.cpp:445-523
RV_HD void combine_digs_digm(ModCtx& m) {
  Cigar& c = *m.cg;
  int soft = c_len(c.op[0]), mch = c_len(c.op[1]);
  int rn = 0;
  while (rn < soft && has_eq(*m.ref, m.position - rn - 1, m.rd->base(soft - rn - 1)) &&
         m.rd->q(soft - rn - 1) > CONF_LOWQUAL)
    rn++;
  if (rn > 0) {
    mch += rn;
    soft -= rn;
    m.position -= rn;
    rn = 0;
  }
  if (soft > 0) {
    BaseSet RN;
    while (rn + 1 < soft && has_eq(*m.ref, m.position - rn - 2, m.rd->base(soft - rn - 2)) &&
           m.rd->q(soft - rn - 2) > CONF_LOWQUAL) {
      rn++;
      if (m.ref->has(m.position - rn - 2)) RN.add(m.ref->at(m.position - rn - 2));
    }
    if ((rn > 4 && RN.size() > 1) || has_eq(*m.ref, m.position - 1, m.rd->base(soft - 1))) {
      mch += rn + 1;
      soft -= rn + 1;
      m.position -= rn + 1;
    }
    if (rn == 0) {
      rn = head_mismatch_walk(m, mch, m.position, soft);
      if (rn > 0 && rn < mch) {
        soft += rn;
        mch -= rn;
        m.position += rn;
      }
    }
  }
  uint32_t nw[2];
  int k = 0;
  if (soft > 0) nw[k++] = c_make(soft, OP_S);
  nw[k++] = c_make(mch, OP_M);
  c.replace(0, 2, nw, k);
}

// head "^(m)M" : combineBeginDigM, cigarModifier.cpp:412-439
RV_HD void combine_begin_digm(ModCtx& m) {
  Cigar& c = *m.cg;
  int mch = c_len(c.op[0]);
  int rn = head_mismatch_walk(m, mch, m.position, 0);
  if (rn > 0 && rn <= 3) {
    mch -= rn;
    uint32_t nw[2] = {c_make(rn, OP_S), c_make(mch, OP_M)};
    c.replace(0, 1, nw, 2);
    m.position += rn;
  }
}

RV_HD bool is_id(int o) { return o == OP_I || o == OP_D; }
RV_HD bool one_digit(int len) { return len >= 0 && len <= 9; }

// shared tail of threeIndels / threeDeletions / twoDeletionsInsertionToComplex (:771-800, :846-858, :907-918)
RV_HD int build_complex(uint32_t* nw, int RDOFF, int dlen, int tslen, int rm, bool three_indels) {
  int k = 0;
  if (tslen <= 0) {
    dlen -= tslen;
    rm += tslen;
    if (three_indels) {
      if (dlen == 0) {
        RDOFF = RDOFF + rm;
        nw[k++] = c_make(RDOFF, OP_M);
        return k;
      } else if (dlen < 0) {
        tslen = -dlen;
        rm += dlen;
        if (rm < 0) {
          RDOFF = RDOFF + rm;
          nw[k++] = c_make(RDOFF, OP_M);
          nw[k++] = c_make(tslen, OP_I);
        } else {
          nw[k++] = c_make(RDOFF, OP_M);
          nw[k++] = c_make(tslen, OP_I);
          nw[k++] = c_make(rm, OP_M);
        }
        return k;
      }
    }
    nw[k++] = c_make(RDOFF, OP_M);
    nw[k++] = c_make(dlen, OP_D);
    nw[k++] = c_make(rm, OP_M);
  } else {
    nw[k++] = c_make(RDOFF, OP_M);
    if (three_indels && dlen == 0) {
      nw[k++] = c_make(tslen, OP_I);
      nw[k++] = c_make(rm, OP_M);
    } else if (three_indels && dlen < 0) {
      rm += dlen;
      nw[k++] = c_make(tslen, OP_I);
      nw[k++] = c_make(rm, OP_M);
    } else {
      nw[k++] = c_make(dlen, OP_D);
      nw[k++] = c_make(tslen, OP_I);
      nw[k++] = c_make(rm, OP_M);
    }
  }
  return k;
}

// Returns true when the CIGAR (or position) was rewritten; `position` is updated only then
// (parseCigar.cpp:559-567).
RV_HDN bool modify_cigar(Cigar& cg, int& position_io, const ReadView& rd, const RefView& ref, int indel,
                         bool* unsupported) {
  Cigar orig;
  orig.n = cg.n;
  orig.overflow = false;
  for (int k = 0; k < cg.n; ++k) orig.op[k] = cg.op[k];
  ModCtx m;
  m.cg = &cg;
  m.position = position_io;
  m.rd = &rd;
  m.ref = &ref;
  m.unsupported = false;
  // :247-256 — a leading D advances a local pointer without shrinking n_cigar (reads one op past the
  // end): refused.  A leading I is rewritten to D.  The two trailing tests compare a whole CIGAR word
  // with an opcode and never fire.
  if (c_op(cg.op[0]) == OP_D) { *unsupported = true; return false; }
  if (c_op(cg.op[0]) == OP_I) cg.op[0] = c_make(c_len(cg.op[0]), OP_D);

  bool flag = true;
  int guard = 0;
  while (flag && indel > 0) {
    if (++guard > 64) { *unsupported = true; break; }
    flag = false;
    Cigar& c = cg;
    // ^(\d+)S(\d+)([ID])  :263-272
    if (c.n >= 2 && c_op(c.op[0]) == OP_S && is_id(c_op(c.op[1]))) {
      int s = c_len(c.op[0]), l = c_len(c.op[1]);
      bool ins = c_op(c.op[1]) == OP_I;
      m.position += ins ? 0 : l;
      uint32_t nw = c_make(s + (ins ? l : 0), OP_S);
      c.replace(0, 2, &nw, 1);
      flag = true;
    }
    // (\d+)([ID])(\d+)S$  :274-279
    if (c.n >= 2 && c_op(c.op[c.n - 1]) == OP_S && is_id(c_op(c.op[c.n - 2]))) {
      int s = c_len(c.op[c.n - 1]), l = c_len(c.op[c.n - 2]);
      bool ins = c_op(c.op[c.n - 2]) == OP_I;
      uint32_t nw = c_make(s + (ins ? l : 0), OP_S);
      c.replace(c.n - 2, 2, &nw, 1);
      flag = true;
    }
    // ^(\d+)S(\d+)M(\d+)([ID])  :281-290
    if (c.n >= 3 && c_op(c.op[0]) == OP_S && c_op(c.op[1]) == OP_M && is_id(c_op(c.op[2]))) {
      int tmid = c_len(c.op[1]);
      if (tmid <= 10) {
        int s = c_len(c.op[0]), l = c_len(c.op[2]);
        bool ins = c_op(c.op[2]) == OP_I;
        m.position += tmid + (ins ? 0 : l);
        uint32_t nw = c_make(s + tmid + (ins ? l : 0), OP_S);
        c.replace(0, 3, &nw, 1);
        flag = true;
      }
    }
    // (\d+)([ID])(\d+)M(\d+)S$  :292-306
    if (c.n >= 3 && c_op(c.op[c.n - 1]) == OP_S && c_op(c.op[c.n - 2]) == OP_M && is_id(c_op(c.op[c.n - 3]))) {
      int tmid = c_len(c.op[c.n - 2]);
      if (tmid <= 10) {
        int s = c_len(c.op[c.n - 1]), l = c_len(c.op[c.n - 3]);
        bool ins = c_op(c.op[c.n - 3]) == OP_I;
        uint32_t nw = c_make(s + tmid + (ins ? l : 0), OP_S);
        c.replace(c.n - 3, 3, &nw, 1);
        flag = true;
      }
    }
    // ^(\d)M(\d+)([ID])(\d+)M  -> beginDigitMNumberIorDNumberM  :309-312, :975-1000
    if (c.n >= 3 && c_op(c.op[0]) == OP_M && one_digit(c_len(c.op[0])) && is_id(c_op(c.op[1])) &&
        c_op(c.op[2]) == OP_M) {
      int tmid = c_len(c.op[0]), l = c_len(c.op[1]), mlen = c_len(c.op[2]);
      bool ins = c_op(c.op[1]) == OP_I;
      int tslen = tmid + (ins ? l : 0);
      m.position += tmid + (ins ? 0 : l);
      int tn = 0;
      while (tn < mlen && has_ne(ref, m.position + tn, rd.base(tslen + tn))) tn++;
      tslen += tn;
      mlen -= tn;
      m.position += tn;
      uint32_t nw[2] = {c_make(tslen, OP_S), c_make(mlen, OP_M)};
      c.replace(0, 3, nw, 2);
      flag = true;
    }
    // (\d+)([ID])(\d)M$  :313-319
    if (c.n >= 2 && c_op(c.op[c.n - 1]) == OP_M && one_digit(c_len(c.op[c.n - 1])) && is_id(c_op(c.op[c.n - 2]))) {
      int tmid = c_len(c.op[c.n - 1]), l = c_len(c.op[c.n - 2]);
      bool ins = c_op(c.op[c.n - 2]) == OP_I;
      uint32_t nw = c_make(tmid + (ins ? l : 0), OP_S);
      c.replace(c.n - 2, 2, &nw, 1);
      flag = true;
    }
    // :321-331 — three searches on the same string, then at most one rewrite.
    //   D_M_D_DD_M_D_I_D_M_D_DD : ^(.*?)(\d+)M(\d+)D(\d+)M(\d+)I(\d+)M(\d+)D(\d+)M
    //   threeDeletionsPattern   : ^(.*?)(\d+)M(\d+)D(\d+)M(\d+)D(\d+)M(\d+)D(\d+)M
    //   threeIndelsPattern      : ^(.*?)(\d+)M(\d+)([DI])(\d+)M(\d+)([DI])(\d+)M(\d+)([DI])(\d+)M
    // The lazy prefix makes each search return its left-most occurrence.
    {
      int i_mdmimdm = -1, i_3del = -1, i_3indel = -1;
      for (int k = 0; k + 7 <= c.n; ++k) {
        bool seven = c_op(c.op[k]) == OP_M && c_op(c.op[k + 2]) == OP_M && c_op(c.op[k + 4]) == OP_M &&
                     c_op(c.op[k + 6]) == OP_M && is_id(c_op(c.op[k + 1])) && is_id(c_op(c.op[k + 3])) &&
                     is_id(c_op(c.op[k + 5]));
        if (!seven) continue;
        int o1 = c_op(c.op[k + 1]), o3 = c_op(c.op[k + 3]), o5 = c_op(c.op[k + 5]);
        if (i_3indel < 0) i_3indel = k;
        if (i_3del < 0 && o1 == OP_D && o3 == OP_D && o5 == OP_D) i_3del = k;
        if (i_mdmimdm < 0 && o1 == OP_D && o3 == OP_I && o5 == OP_D) i_mdmimdm = k;
      }
      int at = -1, kind = 0;
      if (i_mdmimdm >= 0) { at = i_mdmimdm; kind = 1; }
      else if (i_3del >= 0) { at = i_3del; kind = 2; }
      else if (i_3indel >= 0) { at = i_3indel; kind = 3; }
      if (at >= 0) {
        int g[8];
        for (int k = 0; k < 7; ++k) g[k + 1] = c_len(c.op[at + k]);  // g[1]..g[7] = the seven lengths
        int o1 = c_op(c.op[at + 1]), o3 = c_op(c.op[at + 3]), o5 = c_op(c.op[at + 5]);
        int mid = g[3] + g[5];
        int tslen, dlen;
        if (kind == 1) {         // twoDeletionsInsertionToComplex :877-921
          tslen = g[3] + g[4] + g[5];
          dlen = g[2] + g[3] + g[5] + g[6];
        } else if (kind == 2) {  // threeDeletions :815-862
          tslen = g[3] + g[5];
          dlen = g[2] + g[3] + g[4] + g[5] + g[6];
        } else {                 // threeIndels :738-806
          tslen = mid + (o1 == OP_I ? g[2] : 0) + (o3 == OP_I ? g[4] : 0) + (o5 == OP_I ? g[6] : 0);
          dlen = mid + (o1 == OP_D ? g[2] : 0) + (o3 == OP_D ? g[4] : 0) + (o5 == OP_D ? g[6] : 0);
        }
        int refoff = m.position + g[1] + sum_mnd(c, at);
        int rdoff = g[1] + sum_mis(c, at);
        int RDOFF = g[1];
        int rm = g[7];
        int rn = 0;
        while (rdoff + rn < rd.lseq && has_eq(ref, refoff + rn, rd.base(rdoff + rn))) rn++;
        RDOFF += rn;
        dlen -= rn;
        tslen -= rn;
        if (mid <= 15) {
          uint32_t nw[4];
          int k = build_complex(nw, RDOFF, dlen, tslen, rm, kind == 3);
          // the replacement regex is un-anchored and format_first_only: it rewrites the FIRST
          // occurrence of the seven-op shape, which for kinds 2/3 is `at` by construction; for kind 1
          // the prim pattern (M D M I M D M) first occurrence is `at` as well.
          for (int j = 0; j < k; ++j)
            if ((int32_t)(nw[j] >> 4) < 0 || (nw[j] >> 4) > 0x0fffffff) m.unsupported = true;
          c.replace(at, 7, nw, k);
          flag = true;
        }
      }
    }
    // (\d+)D(\d+)M(\d+)([DI])(\d+I)?  -> combineToCloseToCorrect  :333-336, :717-741
    for (int k = 0; k + 3 <= c.n; ++k) {
      if (c_op(c.op[k]) == OP_D && c_op(c.op[k + 1]) == OP_M && is_id(c_op(c.op[k + 2]))) {
        int g1 = c_len(c.op[k]), g2 = c_len(c.op[k + 1]), g3 = c_len(c.op[k + 2]);
        if (g2 <= 15) {
          bool opI = c_op(c.op[k + 2]) == OP_I;
          int dlen = g1 + g2, ilen = g2, used = 3;
          bool trailing_i = k + 3 < c.n && c_op(c.op[k + 3]) == OP_I;
          if (opI) ilen += g3;
          else {
            dlen += g3;
            if (trailing_i) ilen += c_len(c.op[k + 3]);
          }
          if (trailing_i) used = 4;  // the optional group is part of the match that gets replaced
          uint32_t nw[2] = {c_make(dlen, OP_D), c_make(ilen, OP_I)};
          c.replace(k, used, nw, 2);
          flag = true;
        }
        break;  // regex_search only ever sees the left-most occurrence
      }
    }
    // (\D)(\d+)I(\d+)M(\d+)([DI])(\d+I)?  -> combineToCloseToOne  :338-341, :675-705
    for (int k = 1; k + 3 <= c.n; ++k) {
      if (c_op(c.op[k]) == OP_I && c_op(c.op[k + 1]) == OP_M && is_id(c_op(c.op[k + 2]))) {
        int prev = c_op(c.op[k - 1]);
        if (prev != OP_D && prev != OP_H) {
          int g2 = c_len(c.op[k]), g3 = c_len(c.op[k + 1]), g4 = c_len(c.op[k + 2]);
          if (g3 <= 15) {
            bool opI = c_op(c.op[k + 2]) == OP_I;
            int dlen = g3, ilen = g2 + g3, used = 3;
            bool trailing_i = k + 3 < c.n && c_op(c.op[k + 3]) == OP_I;
            if (opI) ilen += g4;
            else {
              dlen += g4;
              if (trailing_i) ilen += c_len(c.op[k + 3]);
            }
            if (trailing_i) used = 4;
            uint32_t nw[2] = {c_make(dlen, OP_D), c_make(ilen, OP_I)};
            c.replace(k, used, nw, 2);
            flag = true;
          }
        }
        break;
      }
    }
    // (\d+)D(\d+)D :342-348 and (\d+)I(\d+)I :349-355 (first occurrence only)
    for (int k = 0; k + 2 <= c.n; ++k)
      if (c_op(c.op[k]) == OP_D && c_op(c.op[k + 1]) == OP_D) {
        uint32_t nw = c_make(c_len(c.op[k]) + c_len(c.op[k + 1]), OP_D);
        c.replace(k, 2, &nw, 1);
        flag = true;
        break;
      }
    for (int k = 0; k + 2 <= c.n; ++k)
      if (c_op(c.op[k]) == OP_I && c_op(c.op[k + 1]) == OP_I) {
        uint32_t nw = c_make(c_len(c.op[k]) + c_len(c.op[k + 1]), OP_I);
        c.replace(k, 2, &nw, 1);
        flag = true;
        break;
      }
  }
  // :365-377
  if (cg.n >= 2 && c_op(cg.op[cg.n - 1]) == OP_S && c_op(cg.op[cg.n - 2]) == OP_M) capture_mis_softly_ms(m);
  else if (cg.n >= 1 && c_op(cg.op[cg.n - 1]) == OP_M) capture_mis_softly_3mm(m);
  if (cg.n >= 2 && c_op(cg.op[0]) == OP_S && c_op(cg.op[1]) == OP_M) combine_digs_digm(m);
  else if (cg.n >= 1 && c_op(cg.op[0]) == OP_M) combine_begin_digm(m);

  if (m.unsupported || cg.overflow) *unsupported = true;
  // :380-401 — "changed" is a comparison of the rendered strings
  bool same = cg.n == orig.n;
  for (int k = 0; same && k < cg.n; ++k) same = cg.op[k] == orig.op[k];
  if (same) return false;
  position_io = m.position;
  return true;
}

// CigarParser::cleanupCigar, parseCigar.cpp:1858-1892
RV_HD void cleanup_cigar(Cigar& c) {
  bool no_match_yet = true;
  for (int k = 0; k < c.n && no_match_yet; ++k) {
    int o = c_op(c.op[k]);
    if (o == OP_I) c.op[k] = c_make(c_len(c.op[k]), OP_S);
    else if (o == OP_H) {}
    else if (o == OP_M || o == OP_EQ || o == OP_X) no_match_yet = false;
  }
  no_match_yet = true;
  for (int k = c.n - 1; k >= 0 && no_match_yet; --k) {
    int o = c_op(c.op[k]);
    if (o == OP_I) c.op[k] = c_make(c_len(c.op[k]), OP_S);
    else if (o == OP_H) {}
    else if (o == OP_M || o == OP_EQ || o == OP_X) no_match_yet = false;
  }
}

// ------------------------------------------------------------------------------------------------
// Fast-path descriptor.  A read whose rewritten CIGAR is [H][S] M [S][H] and whose matched run cannot
// start a multi-nucleotide key (no two mismatches within vext+1 bases, parseCigar.cpp:711-768) contributes
// nothing but independent single-base observations from its M op (:884-937).  Those are not pushed
// through atomics: the walk emits this descriptor and the gather kernel (one lane per reference
// position) accumulates them in registers.
// ------------------------------------------------------------------------------------------------
struct FastDesc {       // 32 bytes
  int32_t m_start;      // reference position of the first matched base
  uint16_t m_len;       // matched bases (== rlen of the read); 0 = no fast-path contribution
  uint16_t rp0;         // read offset of the first matched base
  uint32_t data_off16;  // the read's variable part in the pool
  uint16_t n_cigar;
  int16_t nm;
  uint16_t l_seq;
  uint8_t mapq;
  uint8_t dir;
  uint32_t read_idx;
  uint32_t pad[2];
};

// One observation of the fast path: base/quality of descriptor d at reference position p.
// Returns false when p is outside the matched run or the read base is N (skipped, :686-692).
RV_HD bool fast_obs(const FastDesc& d, int p, const uint8_t* pool, int* allele, int* tp, int* q) {
  int k = p - d.m_start;
  if (k < 0 || k >= (int)d.m_len) return false;
  const uint8_t* var = pool + (size_t)d.data_off16 * 16 + 4 * (size_t)d.n_cigar;
  int r = (int)d.rp0 + k;
  int b = var[r >> 1];
  int nib = (r & 1) ? (b & 15) : (b >> 4);
  int a = nib == 1 ? 0 : nib == 2 ? 1 : nib == 4 ? 2 : nib == 8 ? 3 : -1;
  if (a < 0) return false;
  *allele = a;
  *q = var[((d.l_seq + 1) >> 1) + r];
  *tp = k < (int)d.m_len - k ? k + 1 : (int)d.m_len - k;
  return true;
}

// ------------------------------------------------------------------------------------------------
// CigarParser::parseCigar, parseCigar.cpp:497-974, with its helpers.
// Sink concept:
//   void single(int pos, int allele, bool dir, int tp, int q, int mapq, int nm)   M-path obs, single base key
//   void adj(int pos, int allele, int sign, bool dir, int tp, int q, int mapq, int nm)  addCnt / subtraction
//   void sub_anchor(int pos, int allele, bool dir, int tp, int q, int mapq, int nm)   the anchor base an insertion takes back,
//        parseCigar.cpp:1471-1490: only `if (tv != NULL)`, i.e. if this or an EARLIER read of the region (BAM order) has
//        created the (position, base) entry.  Always true without -T (the read's own M loop has just done it); with -T the
//        anchor can be a trimmed base, so the sinks defer the subtraction until every row's first contributor is known.
//   void cov(int pos)                                                               refCoverage[pos]++
//   void event(const rv_event&)
//   void max_read_len(int tlen)
//   void kept(int aligned_bases) ; void unsupported()
//   int scan_segment(P, rd, ref, m_start, rp, len, indel_follows, SegDesc*)    scan_plain_segment or a faster equivalent:
//                                                              0 = take the per-base walk, 1 = plain, 2 = splittable
//   bool segment(const SegDesc&, bool dir, int mapq, int nm)   a plain matched stretch; a sink that returns false gets
//                                                              the per-base observations
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// Plain segments.  A matched stretch of a read (one M/=/X op, from the bases an indel before it already consumed to
// its end) is PLAIN when the per-base loop of parseCigar (:678-959) can do nothing with it but record one single-base
// observation per base:
//   * every read base is A, C, G or T (an N is skipped without an observation, :686-692; other letters have no row),
//   * the stretch lies inside the loaded reference window,
//   * no two mismatches lie within vext + 1 bases of each other: the multi-nucleotide loop (:711-768) looks at the
//     next base and then up to vext bases further, never beyond the op,
//   * with -k 1 and an I or D op following: no mismatch among the last vext bases (the indel-adjacent complex
//     forms, :779-882).
// Such a stretch leaves a descriptor; the position-major gather kernel accumulates it.  The conditions ignore the
// quality / region tests of the loops they guard, so they are sufficient, not necessary: everything else takes the
// literal per-base walk.
// ------------------------------------------------------------------------------------------------
struct SegDesc {
  int32_t m_start;   // reference position of the first base
  int32_t len;       // bases
  int32_t rp;        // read offset of the first base, soft clips included (index into seq / qual)
  int32_t re;        // ... soft clips excluded: tp of base k = min(re + k + 1, rlen - re - k)
  int32_t rlen;
  uint32_t mm_blocks;        // bit b: a mismatch among bases [16b, 16b + 16) (bit 15: and beyond)
  uint32_t ml[4];            // up to eight mismatches, 16 bits each, oldest in the highest used bits:
                             // 0x8000 | offset << 2 | allele of the read base
  int n_mm;
  int p_first, p_last;       // scan result 2 (splittable): offsets of the first / last base the literal walk must see
};
// Result of a stretch scan: 0 = take the literal walk for all of it, 1 = plain, 2 = plain apart from the bases
// [p_first - (vext + 1), p_last]: what lies before and after that range is plain on its own (split_plain_stretch).
// A base is flagged when it is not A/C/G/T, when it is a mismatch (or such a base) within vext + 1 bases after another
// one — the second element of a close pair; the first lies at most vext + 1 before it — or, with an indel following
// under -k 1, a mismatch among the last vext bases.  The multi-nucleotide loop started at the first element of a pair
// ends at its last flagged element (nothing within vext + 1 after it mismatches), so the literal walk of
// [p_first - (vext + 1), p_last] is self-contained.

RV_HDN int scan_plain_segment(const rv_params& P, const ReadView& rd, const RefView& ref, int m_start, int rp, int len,
                              bool indel_follows, SegDesc* out) {
  if (len <= 0 || len > 8192) return 0;
  if (!ref.has(m_start) || !ref.has(m_start + len - 1)) return 0;
  const int D = P.vext + 1;
  int last_mm = -100000, n_mm = 0, p_first = -1, p_last = -1;
  uint32_t blocks = 0;
  unsigned long long lo = 0, hi = 0;
  for (int k = 0; k < len; ++k) {
    const char b = rd.base(rp + k);
    const int al = allele_of(b);
    const char rc = ref.bases[m_start + k - ref.base_pos];
    if (al < 0 || rc != b) {
      bool flag = al < 0 || k - last_mm <= D || (indel_follows && P.local_realign && len - k <= P.vext);
      last_mm = k;
      if (flag) {
        if (p_first < 0) p_first = k;
        p_last = k;
      }
      if (al >= 0) {
        if (++n_mm > 8) return 0;
        blocks |= 1u << ((k >> 4) < 15 ? (k >> 4) : 15);
        hi = (hi << 16) | (lo >> 48);
        lo = (lo << 16) | (unsigned long long)(0x8000u | ((uint32_t)k << 2) | (uint32_t)al);
      }
    }
  }
  out->mm_blocks = blocks;
  out->ml[0] = (uint32_t)lo; out->ml[1] = (uint32_t)(lo >> 32); out->ml[2] = (uint32_t)hi; out->ml[3] = (uint32_t)(hi >> 32);
  out->n_mm = n_mm;
  out->p_first = p_first;
  out->p_last = p_last;
  if (p_first < 0) return 1;
  return D <= 7 ? 2 : 0;
}

// The mismatch list of the sub-stretch [k0, k1) of a scanned stretch (entries re-based to k0).
RV_HDN void sub_mismatch_list(const SegDesc& full, int k0, int k1, SegDesc* out) {
  unsigned long long lo = (unsigned long long)full.ml[0] | ((unsigned long long)full.ml[1] << 32);
  unsigned long long hi = (unsigned long long)full.ml[2] | ((unsigned long long)full.ml[3] << 32);
  // entries are stored newest (largest offset) in the low bits: collect, then re-insert oldest first
  uint32_t e[8];
  int n = 0;
  for (int j = 0; j < full.n_mm && j < 8; ++j) {
    e[n++] = (uint32_t)lo & 0xffffu;
    lo = (lo >> 16) | (hi << 48);
    hi >>= 16;
  }
  unsigned long long olo = 0, ohi = 0;
  uint32_t blocks = 0;
  int m = 0;
  for (int j = n - 1; j >= 0; --j) {
    const int k = (int)((e[j] >> 2) & 0x1fffu);
    if (k < k0 || k >= k1) continue;
    const int kk = k - k0;
    blocks |= 1u << ((kk >> 4) < 15 ? (kk >> 4) : 15);
    ohi = (ohi << 16) | (olo >> 48);
    olo = (olo << 16) | (unsigned long long)(0x8000u | ((uint32_t)kk << 2) | (e[j] & 3u));
    m++;
  }
  out->mm_blocks = blocks;
  out->ml[0] = (uint32_t)olo; out->ml[1] = (uint32_t)(olo >> 32); out->ml[2] = (uint32_t)ohi; out->ml[3] = (uint32_t)(ohi >> 32);
  out->n_mm = m;
  out->p_first = out->p_last = -1;
}

// ------------------------------------------------------------------------------------------------
// 4-bit reference packing and the nibble-SIMD plain-stretch proof (device: rv_pileup_kernel, rv_walk_kernel; host: the
// single-stepping test tool checks it against scan_plain_segment on every stretch).
// Packed reference: base e of the slice in bits 28 - 4 * (e & 7) of word e >> 3, BAM codes A=1 C=2 G=4 T=8, other=15,
// beyond the slice 0.
// ------------------------------------------------------------------------------------------------
RV_HD uint32_t rv_funnel_l(uint32_t lo, uint32_t hi, int s) {  // upper 32 bits of (hi:lo) << s, 0 <= s < 32
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(lo, hi, s);
#else
  return s ? (hi << s) | (lo >> (32 - s)) : hi;
#endif
}
RV_HD uint32_t rv_funnel_r(uint32_t lo, uint32_t hi, int s) {  // lower 32 bits of (hi:lo) >> s
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, s);
#else
  return s ? (lo >> s) | (hi << (32 - s)) : lo;
#endif
}
RV_HD uint32_t rv_bswap32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(v, 0, 0x0123);
#else
  return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24);
#endif
}
RV_HD int rv_ctz32(uint32_t v) {  // v != 0
#if defined(__CUDA_ARCH__)
  return __ffs((int)v) - 1;
#else
  return __builtin_ctz(v);
#endif
}
RV_HD int rv_clz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __clz((int)v);
#else
  return v ? __builtin_clz(v) : 32;
#endif
}
RV_HD uint32_t pack_ref8(const char* ref, int64_t n, int64_t w) {  // word w of the packed slice
  uint32_t v = 0;
  for (int i = 0; i < 8; ++i) {
    const int64_t e = w * 8 + i;
    uint32_t code = 0;  // beyond the slice: equal to no read base
    if (e < n) {
      const char c = ref[e];
      code = c == 'A' ? 1u : c == 'C' ? 2u : c == 'G' ? 4u : c == 'T' ? 8u : 15u;
    }
    v |= code << (28 - 4 * i);
  }
  return v;
}
RV_HD int nib_allele(int nib) { return (nib >> 1) - (nib >> 3); }  // 1,2,4,8 -> 0,1,2,3

// bit 0 of every nibble = OR of the nibble's four bits
RV_HD uint32_t nib_any(uint32_t x) { return (x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x11111111u; }

// The nibble-SIMD plain-run proof.  The matched stretch [rp0, rp0 + ml) of a read is compared with the reference 8
// bases per step: BAM's 4-bit bases XOR the 4-bit reference (funnel-shifted to the read's phase) give one mismatch flag
// per nibble; shifted copies of the flag word, carried across words, prove that no two mismatches lie within D = vext + 1
// bases, i.e. that no base of the stretch can start a multi-nucleotide key (parseCigar.cpp:711-768).  Read bases other
// than A, C, G, T make the stretch not plain.  Returns bad == 0 (more than eight mismatches are left to the caller);
// mismatches are listed as 0x8000 | offset << 2 | allele, 16 bits each, the newest in the low bits.
// sq: the read's packed bases (word 0 = bases 0..7); E0: reference-slice index of read base 0 (>= 0).
struct PlainScan {
  uint32_t mm_blocks;
  unsigned long long ml_lo, ml_hi;
  int ml_n;
  uint32_t special;  // a read base that is not A, C, G, T was seen
  int p_first, p_last;  // offsets of the first / last flagged base (see scan_plain_segment), -1 = none
};
RV_HD bool simd_plain_scan(const uint32_t* sq, const uint32_t* ref4, int E0, int rp0, int ml, int D,
                                                PlainScan* out) {
  const int w_first = rp0 >> 3, w_last = (rp0 + ml - 1) >> 3;
  const int e = E0 + 8 * w_first;
  const int sh = (e & 7) * 4;
  const uint32_t* rw = ref4 + (e >> 3);
  uint32_t r_lo = rw[0];
  uint32_t prev = 0, bad = 0, seen = 0, spec = 0, mm_blocks = 0;
  unsigned long long ml_lo = 0, ml_hi = 0;
  int ml_n = 0, p_first = -1, p_last = -1;
  for (int wi = w_first; wi <= w_last; ++wi) {
    const uint32_t r_hi = *++rw;
    const uint32_t rf = rv_funnel_l(r_hi, r_lo, sh);
    r_lo = r_hi;
    const uint32_t b8 = rv_bswap32(sq[wi]);  // base 8*wi in the top nibble
    uint32_t vm = 0xffffffffu;
    if (wi == w_first) vm >>= 4 * (rp0 & 7);
    if (wi == w_last) {
      const int hi = ((rp0 + ml - 1) & 7) + 1;
      if (hi < 8) vm &= ~(0xffffffffu >> (4 * hi));
    }
    const uint32_t nz = nib_any(b8 ^ rf) & vm;
    // read bases other than A, C, G, T (N included): zero nibble, or more than one bit in the nibble
    // (the nibble-wise x & (x - 1) borrows out of a zero nibble: nibbles outside the stretch are forced to 1 first, so
    // that a zero there — padding, or the byte that follows the packed bases — cannot taint its neighbour)
    const uint32_t b8s = (b8 & vm) | (0x11111111u & ~vm);
    const uint32_t special = (~nib_any(b8s) | nib_any(b8s & (b8s - 0x11111111u))) & 0x11111111u & vm;
    // close pairs: a mismatch — or a base that is not A/C/G/T, which the multi-nucleotide loop would take in as well —
    // within D bases after another one
    const uint32_t nzs = nz | special;
    uint32_t near;
    if (D == 3) {
      near = nzs & (rv_funnel_r(nzs, prev, 4) | rv_funnel_r(nzs, prev, 8) | rv_funnel_r(nzs, prev, 12));
    } else if (D <= 7) {
      near = 0;
      for (int d = 1; d <= D; ++d) near |= nzs & rv_funnel_r(nzs, prev, 4 * d);
    } else {
      near = (nzs & (nzs - 1)) | ((nzs && seen) ? 1u : 0u);  // conservative: any two mismatches in the stretch
      seen |= nzs;
    }
    const uint32_t flagged = near | special;
    if (flagged) {  // base 8*wi + i has its flag at bit 28 - 4i
      const int f_lo = 8 * wi + (rv_clz32(flagged) >> 2) - rp0;
      const int f_hi = 8 * wi + 7 - (rv_ctz32(flagged) >> 2) - rp0;
      if (p_first < 0) p_first = f_lo;
      p_last = f_hi;
    }
    bad |= flagged;
    spec |= special;
    prev = nzs;
    if (nz & ~special) {  // 16-base blocks of the stretch this word's mismatches may lie in (a word touches at most two)
      const int k_lo = 8 * wi - rp0 > 0 ? 8 * wi - rp0 : 0;
      const int k_hi = 8 * wi + 7 - rp0 < ml - 1 ? 8 * wi + 7 - rp0 : ml - 1;
      mm_blocks |= (1u << ((k_lo >> 4) < 15 ? (k_lo >> 4) : 15)) | (1u << ((k_hi >> 4) < 15 ? (k_hi >> 4) : 15));
      for (uint32_t z = nz & ~special; z;) {  // the A/C/G/T mismatches themselves: base 8*wi + i has its flag at bit 28 - 4i
        const int i = rv_clz32(z) >> 2;
        z &= ~(0x10000000u >> (4 * i));
        const uint32_t en = 0x8000u | ((uint32_t)(8 * wi + i - rp0) << 2) | ((uint32_t)nib_allele((b8 >> (28 - 4 * i)) & 15u) & 3u);
        ml_hi = (ml_hi << 16) | (ml_lo >> 48);
        ml_lo = (ml_lo << 16) | en;
        ml_n++;
      }
    }
  }
  out->mm_blocks = mm_blocks;
  out->ml_lo = ml_lo;
  out->ml_hi = ml_hi;
  out->ml_n = ml_n;
  out->special = spec;
  out->p_first = p_first;
  out->p_last = p_last;
  return bad == 0;
}
// is any listed mismatch among the first `head` or the last `tail` bases of a stretch of ml bases?
RV_HD bool mismatch_near_ends(const PlainScan& ps, int ml, int head, int tail) {
  unsigned long long lo = ps.ml_lo, hi = ps.ml_hi;
  const int n = ps.ml_n < 8 ? ps.ml_n : 8;
  for (int j = 0; j < n; ++j) {
    const int k = (int)((lo >> 2) & 0x1fffu);
    if (k < head || k >= ml - tail) return true;
    lo = (lo >> 16) | (hi << 48);
    hi >>= 16;
  }
  return ps.ml_n > 8;  // more mismatches than the list holds: unknown, say yes
}


// scan_plain_segment's verdict from the nibble-SIMD proof (the caller has checked the window and E0 >= 0)
RV_HD int simd_scan_kind(const rv_params& P, const uint32_t* sq, const uint32_t* ref4, int E0, int rp, int len, bool indel_follows,
                         PlainScan* ps) {
  const bool clean = simd_plain_scan(sq, ref4, E0, rp, len, P.vext + 1, ps);
  if (ps->ml_n > 8) return 0;
  if (indel_follows && P.local_realign) {  // mismatches among the last vext bases are flagged as well
    unsigned long long lo = ps->ml_lo, hi = ps->ml_hi;
    for (int j = 0; j < ps->ml_n; ++j) {
      const int k = (int)((lo >> 2) & 0x1fffu);
      if (len - k <= P.vext) {
        if (ps->p_first < 0 || k < ps->p_first) ps->p_first = k;
        if (k > ps->p_last) ps->p_last = k;
      }
      lo = (lo >> 16) | (hi << 48);
      hi >>= 16;
    }
  }
  if (clean && ps->p_first < 0) return 1;
  return P.vext + 1 <= 7 ? 2 : 0;
}

struct WalkState {
  int start, rp, re, offset, clen;  // start, readPositionIncluding/ExcludingSoftClipped, offset, cigar_element_length
  int seq_no;
};

template <class Sink>
RV_HD void emit_event(Sink& s, WalkState& w, int region_idx, uint32_t read_idx, int kind, int pos, const Key* key,
                      int flags, bool dir, int tp, int qsum, int qcnt, int mapq, int nm, int aux0, int aux1, int aux2) {
  rv_event e;
  e.region = region_idx;
  e.pos = pos;
  e.read_idx = read_idx;
  e.seq_no = (uint16_t)w.seq_no++;
  e.kind = (uint8_t)kind;
  e.flags = (uint8_t)(flags | ((key && key->trunc) ? RV_EVF_KEY_TRUNC : 0));
  e.tp = tp;
  e.nm = nm;
  e.qsum = qsum;
  e.qcnt = qcnt;
  e.dir = dir ? 1 : 0;
  e.mapq = (uint8_t)mapq;
  e.keylen = (uint8_t)(key ? key->n : 0);
  e.pad = 0;
  e.aux0 = aux0;
  e.aux1 = aux1;
  e.aux2 = aux2;
  // (whole words: the key area is 4-byte aligned in both structs; bytes past the key are zero)
  const int kn = key ? key->n : 0;
  for (int k = 0; k < RV_EVENT_KEY_MAX; k += 4) {
    uint32_t w = 0;
    if (k < kn) {
      w = *(const uint32_t*)(key->c + k);  // (little endian on both sides; bytes past the key's end are masked off)
      if (kn - k < 4) w &= (1u << (8 * (kn - k))) - 1u;
    }
    *(uint32_t*)(e.key + k) = w;
  }
  if (key && key->trunc) s.unsupported();  // counted, never scored silently (the host stage refuses the batch)
  s.event(e);
}

RV_HD bool key_is_mnp(const Key& k) {  // isBEGIN_ATGC_AMP_ATGCs_END, parseCigar.cpp:480-495
  if (k.n > 2 && k.c[1] == '&' && is_atgc(k.c[0])) {
    for (int i = 2; i < k.n; ++i) if (!is_atgc(k.c[i])) return false;
    return true;
  }
  return false;
}

// the look-ahead scan shared by findOffset (:208-256) and the in-line copies in process_insertion
// (:1367-1389) / process_deletion (:1576-1595,:1622-1644,:1659-1680).  `n_break_ref` selects the variant
// that also stops on a reference 'N'.  Returns offset, writes nmoff increment.
RV_HD int lookahead_offset(const rv_params& P, const ReadView& rd, const RefView& ref, int ref_pos, int read_pos,
                           int mlen, int n_break_mode, int* tnm) {
  int offset = 0, vsn = 0;
  *tnm = 0;
  for (int vi = 0; vsn <= P.vext && vi < mlen; vi++) {
    char ch = rd.base(read_pos + vi);
    if (ch == 'N') break;
    if (rd.q(read_pos + vi) < iceil(P.goodq)) break;  // integer q: q < g <=> q < ceil(g)
    if (n_break_mode == 1 && has_eq(ref, ref_pos + vi, 'N')) break;  // :1582 (before the has() test)
    if (ref.has(ref_pos + vi)) {
      char rc = ref.at(ref_pos + vi);
      if (n_break_mode == 2 && rc == 'N') break;                     // :1635, :1671
      if (ch != rc) { offset = vi + 1; (*tnm)++; vsn = 0; }
      else vsn++;
    }
  }
  return offset;
}

// The -t filter of RecordPreprocessor::next_record (recordPreprocessor.cpp:153-176) without its running set.
// The reference keeps a set of keys that it clears whenever a record's start differs from the start of the last
// record it inserted; every key begins with the record's start and a BAM is coordinate-sorted, so the set only ever
// holds keys of the current start.  "Seen before" is therefore: an EARLIER record of the same fetch (same region,
// passing the htslib overlap test and the filters that run before the -t block, :121-146) with the same start and
// the same key.  Two key kinds, in this order (:157, :166):
//   mate start < 10                  -> POS - mate name - PNEXT   (mate name: "*" unpaired, "=" same contig, else name)
//   else paired and flagged unmapped -> POS - CIGAR string
// A key of one kind never equals a key of the other (a CIGAR string holds no '-').  One thread decides for its own
// read by scanning back over the reads that share its start; eligible reads are rare (mate start < 10, or 0x4 reads
// that survive -F), so the scan is off the common path.  `reads` is indexable by batch read index.
RV_HDN bool record_prefilters_pass(const rv_params& P, const rv_region& R, const rv_read& r) {
  if (!(r.pos - 1 < R.end && r.end_pos > R.start - 1)) return false;  // never returned by the iterator
  if ((r.flag & P.samfilter) != 0) return false;
  if ((int)r.mapq < P.mapping_quality) return false;
  if (r.l_seq == 1) return false;
  return true;
}

RV_HDN int dedup_key_kind(const rv_read& r) {
  if (r.mpos < 10) return 1;
  if ((r.flag & 1) && (r.flag & 4)) return 2;
  return 0;
}

RV_HDN bool is_duplicate_read(const rv_params& P, const rv_region& R, const rv_read* reads, const uint8_t* pool,
                              int64_t read_idx) {
  const rv_read me = reads[read_idx];
  const int kind = dedup_key_kind(me);
  if (kind == 0 || !record_prefilters_pass(P, R, me)) return false;
  const uint32_t* my_cigar = (const uint32_t*)(pool + (size_t)me.data_off16 * 16);
  for (int64_t j = read_idx - 1; j >= R.read_lo; --j) {
    const rv_read o = reads[j];
    if (o.pos != me.pos) break;
    if (dedup_key_kind(o) != kind || !record_prefilters_pass(P, R, o)) continue;
    if (kind == 1) {
      if (o.mpos != me.mpos) continue;
      const bool mp = (me.flag & 1) != 0, op = (o.flag & 1) != 0;
      if (mp != op) continue;
      if (mp && (me.mate_same_tid != o.mate_same_tid || (!me.mate_same_tid && me.mtid != o.mtid))) continue;
      return true;
    }
    if (o.n_cigar != me.n_cigar) continue;
    const uint32_t* oc = (const uint32_t*)(pool + (size_t)o.data_off16 * 16);
    bool same = true;
    for (int k = 0; k < (int)me.n_cigar; ++k) same = same && oc[k] == my_cigar[k];
    if (same) return true;
  }
  return false;
}

// Everything parseCigar does before its op loop (:497-625): filters, CIGAR rewrite, clean-up, lengths.
struct Prep {
  Cigar cg;
  int n_cigar, nm, position, rlen, tlen, mapq;
  bool dir, ok, fast_shape;
  int m_start, m_len, rp0;  // the single matched run of a fast-shaped read
  int rec_ref_len;          // getReferenceLength(record) as skipOverlappingReads sees it (--UN only)
};

template <class Sink>
RV_HDN void prepare_read(const rv_params& P, const rv_region& R, const rv_read& rdh, const uint8_t* pool,
                         const RefView& ref, Sink& sink, bool want_fast, Prep& pr) {
  pr.ok = false;
  pr.fast_shape = false;
  Cigar& cg = pr.cg;
  const uint8_t* var = pool + (size_t)rdh.data_off16 * 16;
  const uint32_t* cig_in = (const uint32_t*)var;
  ReadView rd;
  rd.seq4 = var + 4 * (size_t)rdh.n_cigar;
  rd.lseq = rdh.l_seq;
  rd.qual = rd.seq4 + ((rdh.l_seq + 1) >> 1);

  // ---- RecordPreprocessor::next_record filters, recordPreprocessor.cpp:121-146 -------------------
  if ((rdh.flag & P.samfilter) != 0) return;
  if ((int)rdh.mapq < P.mapping_quality) return;
  if (rdh.l_seq == 1) return;
  // ---- parseCigar.cpp:505-534 ---------------------------------------------------------------------
  int n_cigar = rdh.n_cigar;
  if (n_cigar <= 0) return;
  if (n_cigar > RV_MAX_OPS - 8) { sink.unsupported(); return; }
  cg.n = n_cigar;
  cg.overflow = false;
  int indel = 0;
  for (int k = 0; k < n_cigar; ++k) {
    cg.op[k] = cig_in[k];
    int o = c_op(cig_in[k]);
    if (o == OP_I || o == OP_D) indel += c_len(cig_in[k]);
    if (o == OP_P || o == OP_B) { sink.unsupported(); return; }  // Appendix A-23
  }
  int nm = 0;
  if (rdh.nm >= 0) {
    nm = (int)rdh.nm - indel;
    if (nm > P.mismatch) return;
  } else {
    if (rdh.flag & 4) return;
    nm = 0;
  }
  const bool dir = (rdh.flag & 16) != 0;
  const int mapq = rdh.mapq;
  int position = rdh.pos;
  // ---- CigarModifier, parseCigar.cpp:547-574 -------------------------------------------------------
  if (P.local_realign) {
    bool unsup = false;
    modify_cigar(cg, position, rd, ref, indel, &unsup);
    if (unsup) { sink.unsupported(); return; }
  }
  cleanup_cigar(cg);
  n_cigar = cg.n;
  // --UN (isReadsOverlap :196-206) asks the RECORD for its reference length and end.  CigarModifier writes the
  // rewritten ops over the record's own CIGAR (cigarstr_2cigar, cigarModifier.cpp:386) while core.n_cigar and
  // core.pos keep their values: the record then answers with the first core.n_cigar entries of that array — the
  // rewritten ops followed by whatever original ops they did not reach.
  pr.rec_ref_len = 0;
  if (P.uniq_un) {
    int rl = 0;
    for (int k = 0; k < (int)rdh.n_cigar; ++k) {
      const uint32_t e = k < n_cigar ? cg.op[k] : cig_in[k];
      const int o = c_op(e);
      if (o == OP_M || o == OP_D || o == OP_N || o == OP_EQ || o == OP_X) rl += c_len(e);
    }
    pr.rec_ref_len = rl;
  }
  // :584-588
  if (c_op(cg.op[0]) == OP_S && c_len(cg.op[0]) >= 10 && c_op(cg.op[n_cigar - 1]) == OP_S &&
      c_len(cg.op[n_cigar - 1]) >= 10)
    return;
  // :591-601 (only M and I count; '=' / 'X' do not — literal)
  int rlen = 0, tlen = 0, aligned = 0;
  for (int k = 0; k < n_cigar; ++k) {
    int o = c_op(cg.op[k]), l = c_len(cg.op[k]);
    if (o == OP_M || o == OP_I) rlen += l;
    if (o == OP_M || o == OP_I || o == OP_S) tlen += l;
    if (o == OP_M || o == OP_EQ || o == OP_X) aligned += l;
  }
  if (P.minmatch != 0 && rlen < P.minmatch) return;
  sink.max_read_len(tlen);
  if (rdh.flag & 2048) return;  // :603
  sink.kept(aligned);
  // fast-path shape: [H][S] M [S][H], no trimming option.  The matched run starts at `position`
  // (process_softclip resets start to it, :1300) and at the read offset after the leading clips.
  bool fast_shape = want_fast && P.trim_bases_after == 0 && tlen < 65536;
  int m_len = 0, rp0 = 0;
  if (fast_shape) {
    int n_m = 0;
    for (int k = 0; k < n_cigar; ++k) {
      int o = c_op(cg.op[k]);
      if (o == OP_M) { n_m++; m_len = c_len(cg.op[k]); }
      else if (o == OP_S) { if (n_m == 0) rp0 += c_len(cg.op[k]); }
      else if (o != OP_H) fast_shape = false;
    }
    if (n_m != 1) fast_shape = false;
  }
  pr.n_cigar = n_cigar;
  pr.nm = nm;
  pr.position = position;
  pr.rlen = rlen;
  pr.tlen = tlen;
  pr.mapq = mapq;
  pr.dir = dir;
  pr.fast_shape = fast_shape;
  pr.m_start = position;
  pr.m_len = m_len;
  pr.rp0 = rp0;
  pr.ok = true;
}

// plain_hint: 1 = the matched run was already proven plain (warp-cooperative scan in the kernel),
// 0 = proven not plain, -1 = decide here with the scalar scan.
// has_work: on the device ALL lanes of a warp call walk_read together; a lane without a read passes false.
template <class Sink>
RV_HDN void walk_read(const rv_params& P, const rv_region& R, int region_idx, const rv_read& rdh, const uint8_t* pool,
                      const RefView& ref, uint32_t read_idx, Sink& sink, Prep& pr, FastDesc* fast, int plain_hint,
                      bool has_work = true) {
  Cigar& cg = pr.cg;
  const uint8_t* var = pool + (size_t)rdh.data_off16 * 16;
  ReadView rd;
  rd.seq4 = var + 4 * (size_t)rdh.n_cigar;
  rd.lseq = rdh.l_seq;
  rd.qual = rd.seq4 + ((rdh.l_seq + 1) >> 1);
  const int n_cigar = pr.n_cigar, nm = pr.nm, position = pr.position, rlen = pr.rlen, tlen = pr.tlen, mapq = pr.mapq;
  const bool dir = pr.dir;
  const bool fast_shape = pr.fast_shape && fast != (FastDesc*)0 && plain_hint != 0;
  // qualities are integers: comparisons with the double thresholds are done on their ceilings
  int goodq_i = iceil(P.goodq), goodq5_i = iceil(P.goodq + 5);
  int trim_after = P.trim_bases_after, vext = P.vext, r_start = R.start, r_end = R.end;
  RV_KEEP_REG(goodq_i); RV_KEEP_REG(goodq5_i); RV_KEEP_REG(trim_after); RV_KEEP_REG(vext); RV_KEEP_REG(r_start); RV_KEEP_REG(r_end);
  WalkState w;
  w.start = position;
  w.offset = 0;
  w.rp = 0;
  w.re = 0;
  w.seq_no = 0;
  const int mate_start = rdh.mpos;
  const bool paired_same = (rdh.flag & 1) && rdh.mate_same_tid;

  bool need_break = true;
  int ci = 0;
  // One CIGAR op per round.  On the device the rounds are warp-synchronous: the 32 lanes (32 reads) enter every round
  // together, so lanes whose op is of the same kind run its loops side by side instead of one after the other
  // (independent thread scheduling never brings lanes back together inside a loop they entered at different times).
  auto op_step = [&]() -> bool {  // true = the read is finished
    // :630-634 / skipOverlappingReads :182-206 — only evaluated before the first op
    if (need_break) {
      bool skip = false;
      if (P.uniq_u && paired_same && !dir && w.start >= mate_start) skip = true;
      if (!skip && P.uniq_un && (rdh.flag & 1) && paired_same) {
        // isReadsOverlap :196-206 ; getReferenceLength / getAlignmentEnd of the record (see prepare_read)
        const int ref_len = pr.rec_ref_len;
        const int rec_end = rdh.pos - 1 + (ref_len ? ref_len : 1);  // bam_endpos
        bool ov;
        if (position >= mate_start) ov = w.start >= mate_start && w.start <= mate_start + ref_len - 1;
        else ov = w.start >= mate_start && mate_start <= rec_end;
        if (ov) skip = true;
      }
      if (skip) return true;
    }
    need_break = false;
    int c_operator = c_op(cg.op[ci]);
    w.clen = c_len(cg.op[ci]);
    if ((ci == 0 || ci == n_cigar - 1) && c_operator == OP_I) c_operator = OP_S;

    if (c_operator == OP_N) {  // processNotMatched :1733-1753 — splice bookkeeping refused (A-16)
      sink.unsupported();
      w.start += w.clen;
      w.offset = 0;
      return false;
    }
    if (c_operator == OP_S) {
      // ---- process_softclip :1112-1301 -----------------------------------------------------------
      if (ci == 0) {
        while (w.clen - 1 >= 0 && w.start - 1 > 0 && w.start - 1 <= R.chr_len &&
               has_eq(ref, w.start - 1, rd.base(w.clen - 1)) && rd.q(w.clen - 1) > 10) {
          char rc = ref.at(w.start - 1);
          int al = allele_of(rc);
          if (al >= 0) sink.adj(w.start - 1, al, +1, dir, w.clen, rd.q(w.clen - 1), mapq, nm);
          else sink.unsupported();
          sink.cov(w.start - 1);
          w.start--;
          w.clen--;
        }
        if (w.clen > 0) {
          int qsum = 0, nhi = 0, nlo = 0;
          for (int si = w.clen - 1; si >= 0; si--) {
            if (rd.base(si) == 'N') break;
            int bq = rd.q(si);
            if (bq <= 12) nlo++;
            if (nlo > 1) break;
            qsum += bq;
            nhi++;
          }
          if (nhi >= 1 && nhi > nlo && w.start >= R.start && w.start <= R.end)
            emit_event(sink, w, region_idx, read_idx, RV_EV_SC5, w.start, (const Key*)0, 0, dir, w.clen, qsum, nhi,
                       mapq, nm, w.clen, nhi, w.clen - 1);
        }
        w.clen = c_len(cg.op[ci]);
      } else if (ci == n_cigar - 1) {
        while (w.rp < rdh.l_seq && has_eq(ref, w.start, rd.base(w.rp)) && rd.q(w.rp) > 10) {
          char rc = ref.at(w.start);
          int al = allele_of(rc);
          if (al >= 0) sink.adj(w.start, al, +1, dir, tlen - w.re, rd.q(w.rp), mapq, nm);
          else sink.unsupported();
          sink.cov(w.start);
          w.rp++;
          w.start++;
          w.clen--;
          w.re++;
        }
        if (w.rp < rdh.l_seq) {
          int qsum = 0, nhi = 0, nlo = 0;
          for (int si = 0; si < w.clen; si++) {
            if (rd.base(w.rp + si) == 'N') break;
            int bq = rd.q(w.rp + si);
            if (bq <= 12) nlo++;
            if (nlo > 1) break;
            qsum += bq;
            nhi++;
          }
          if (nhi >= 1 && nhi > nlo && w.start >= R.start && w.start <= R.end)
            emit_event(sink, w, region_idx, read_idx, RV_EV_SC3, w.start, (const Key*)0, 0, dir, w.clen, qsum, nhi,
                       mapq, nm, w.clen, nhi, w.rp);
        }
      }
      w.rp += w.clen;
      w.offset = 0;
      w.start = position;
      return false;
    }
    if (c_operator == OP_H) { w.offset = 0; return false; }

    if (c_operator == OP_I) {
      // ---- process_insertion :1303-1523 ----------------------------------------------------------
      w.offset = 0;
      if ((n_cigar > ci + 1 && c_op(cg.op[ci + 1]) == OP_N) || (ci > 0 && c_op(cg.op[ci - 1]) == OP_N)) {
        w.rp += w.clen;
        return false;
      }
      Key key;
      key.clear();
      for (int k = 0; k < w.clen; ++k) key.push(rd.base(w.rp + k));
      int qsum = 0, qcnt = 0;
      for (int k = w.rp; k < w.rp + w.clen; ++k) qsum += rd.q(k);
      qcnt += w.clen;
      Key ss;
      ss.clear();
      int multoffs = 0, multoffp = 0, nmoff = 0;
      // isInsertionOrDeletionWithNextMatched :1816-1824 (reads cigar[ci+3] with only ci+2 guaranteed: A-6)
      bool combo = P.local_realign && n_cigar > ci + 2 && c_len(cg.op[ci + 1]) <= P.vext &&
                   c_op(cg.op[ci + 1]) == OP_M && is_id(c_op(cg.op[ci + 2]));
      if (combo) {
        if (ci + 3 >= n_cigar) { sink.unsupported(); return true; }
        combo = !is_id(c_op(cg.op[ci + 3]));
      }
      if (combo) {
        int mLen = c_len(cg.op[ci + 1]), indelLen = c_len(cg.op[ci + 2]);
        int begin = w.rp + w.clen;
        bool next_ins = c_op(cg.op[ci + 2]) == OP_I;
        // appendSegments(..., isInsertion = true) :1756-1815
        key.push('#');
        for (int k = 0; k < mLen; ++k) { key.push(rd.base(begin + k)); qsum += rd.q(begin + k); }
        key.push('^');
        if (next_ins) {
          for (int k = 0; k < indelLen; ++k) { key.push(rd.base(begin + mLen + k)); qsum += rd.q(begin + mLen + k); }
          qcnt += indelLen;
        } else {
          key.push_int(indelLen);
          qsum += rd.q(begin + mLen);
          qcnt += 1;
        }
        multoffs += mLen + (next_ins ? 0 : indelLen);
        multoffp += mLen + (next_ins ? indelLen : 0);
        int ci6 = n_cigar > ci + 3 ? c_len(cg.op[ci + 3]) : 0;
        if (ci6 != 0 && c_op(cg.op[ci + 3]) == OP_M) {
          // findOffset :208-256 (coverage of the offset bases is incremented without a region test)
          int tnm;
          int rpos = w.start + multoffs, qpos = w.rp + w.clen + multoffp;
          int off = lookahead_offset(P, rd, ref, rpos, qpos, ci6, 0, &tnm);
          if (off > 0) {
            for (int k = 0; k < off; ++k) { ss.push(rd.base(qpos + k)); qsum += rd.q(qpos + k); sink.cov(rpos + k); }
          }
          w.offset = off;
          qcnt += off;
          // note: findOffset's mismatch count is NOT added to nmoff on this path (:1348-1355)
        }
        ci += 2;
      } else if (P.local_realign && n_cigar > ci + 1 && c_op(cg.op[ci + 1]) == OP_M) {  // isNextMatched :1837
        int tnm;
        int qpos = w.rp + w.clen;
        int off = lookahead_offset(P, rd, ref, w.start, qpos, c_len(cg.op[ci + 1]), 0, &tnm);
        w.offset = off;
        nmoff += tnm;
        if (off != 0) {
          for (int k = 0; k < off; ++k) { ss.push(rd.base(qpos + k)); qsum += rd.q(qpos + k); sink.cov(w.start + k); }
          qcnt += off;
        }
      }
      if (w.offset > 0) {
        key.push('&');
        for (int k = 0; k < ss.n; ++k) key.push(ss.c[k]);
      }
      if (w.start - 1 >= R.start && w.start - 1 <= R.end && !key.contains('N')) {
        int inspos = w.start - 1;
        bool all_atgc = key.n > 0;
        for (int k = 0; k < key.n; ++k) all_atgc = all_atgc && is_atgc(key.c[k]);
        if (all_atgc) {
          // adjInsPos, include/VariationUtils.h:99-113
          int n = 1, len = key.n, bi = w.start - 1;
          while (ref.at(bi) == key.c[len - n]) {
            n++;
            if (n > len) n = 1;
            bi--;
          }
          if (w.rp - 1 - (w.start - 1 - bi) > 0) {
            inspos = bi;
            if (n > 1) {  // rotate: last n-1 chars first
              Key t;
              t.clear();
              for (int k = len - (n - 1); k < len; ++k) t.push(key.c[k]);
              for (int k = 0; k < len - (n - 1); ++k) t.push(key.c[k]);
              key = t;
            }
          }
        }
        key.prepend("+", 1);
        int tp = w.re < rlen - w.re ? w.re + 1 : rlen - w.re;
        emit_event(sink, w, region_idx, read_idx, RV_EV_IN, inspos, &key, RV_EVF_PINS, dir, tp, qsum, qcnt, mapq,
                   nm - nmoff, 0, 0, 0);
        // :1471-1490 — take the anchor base's observation back from the reference allele
        int index = w.rp - 1 - (w.start - 1 - inspos);
        if (inspos > position && has_eq(ref, inspos, rd.base(index))) {
          int al = allele_of(rd.base(index));
          if (al >= 0) sink.sub_anchor(inspos, al, dir, tp, rd.q(index), mapq, nm - nmoff);
          else sink.unsupported();
        }
        // :1497-1515 — insertion right behind a leading S/H: one extra reference observation
        if (ci == 1 && (c_op(cg.op[0]) == OP_S || c_op(cg.op[0]) == OP_H)) {
          Key rk;
          rk.clear();
          rk.push(ref.at(inspos));
          emit_event(sink, w, region_idx, read_idx, RV_EV_TTREF, inspos, &rk, 0, dir, tp, qsum, qcnt, mapq,
                     nm - nmoff, 0, 0, 0);
          sink.cov(inspos);
        }
      }
      w.rp += w.clen + w.offset + multoffp;
      w.re += w.clen + w.offset + multoffp;
      w.start += w.offset + multoffs;
      return false;
    }

    if (c_operator == OP_D) {
      // ---- process_deletion :1525-1727 -----------------------------------------------------------
      w.offset = 0;
      if (ci + 1 >= n_cigar) { sink.unsupported(); return true; }  // reads cigar[ci+1] unconditionally (A-6)
      if (c_op(cg.op[ci + 1]) == OP_N || (ci > 1 && c_op(cg.op[ci - 1]) == OP_N)) {
        w.rp += w.clen;
        return false;
      }
      Key key;
      key.clear();
      key.push('-');
      key.push_int(w.clen);
      Key app;
      app.clear();
      int q_before = rd.q(w.rp - 1);
      int qsum = 0, qcnt = 0;
      int multoffs = 0, multoffp = 0, nmoff = 0;
      bool combo = P.local_realign && n_cigar > ci + 2 && c_len(cg.op[ci + 1]) <= P.vext &&
                   c_op(cg.op[ci + 1]) == OP_M && is_id(c_op(cg.op[ci + 2]));
      if (combo) {
        if (ci + 3 >= n_cigar) { sink.unsupported(); return true; }
        combo = !is_id(c_op(cg.op[ci + 3]));
      }
      if (combo) {
        int mLen = c_len(cg.op[ci + 1]), indelLen = c_len(cg.op[ci + 2]);
        int begin = w.rp;
        bool next_ins = c_op(cg.op[ci + 2]) == OP_I;
        // appendSegments(..., isInsertion = false)
        key.push('#');
        for (int k = 0; k < mLen; ++k) { key.push(rd.base(begin + k)); qsum += rd.q(begin + k); }
        key.push('^');
        if (next_ins) {
          for (int k = 0; k < indelLen; ++k) { key.push(rd.base(begin + mLen + k)); qsum += rd.q(begin + mLen + k); }
          qcnt += indelLen;
        } else {
          key.push_int(indelLen);
        }
        multoffs += mLen + (next_ins ? 0 : indelLen);
        multoffp += mLen + (next_ins ? indelLen : 0);
        // isNextAfterNumMatched(ci, 3): n_cigar > ci+3 && op[ci+1] == M  (:1084-1087 tests ci+1, literal)
        if (n_cigar > ci + 3 && c_op(cg.op[ci + 1]) == OP_M) {
          int tn = w.rp + multoffp, ts = w.start + multoffs + w.clen;
          int tnm;
          int off = lookahead_offset(P, rd, ref, ts, tn, c_len(cg.op[ci + 3]), 1, &tnm);
          w.offset = off;
          nmoff += tnm;
          if (off != 0) {
            for (int k = 0; k < off; ++k) { app.push(rd.base(tn + k)); qsum += rd.q(tn + k); }
            qcnt += off;
          }
        }
        ci += 2;
      } else if (P.local_realign && n_cigar > ci + 1 && c_op(cg.op[ci + 1]) == OP_I) {  // isNextInsertion :163
        int insLen = c_len(cg.op[ci + 1]);
        key.push('^');
        for (int k = 0; k < insLen; ++k) { key.push(rd.base(w.rp + k)); qsum += rd.q(w.rp + k); }
        qcnt += insLen;
        multoffp += insLen;
        if (n_cigar > ci + 2 && c_op(cg.op[ci + 1]) == OP_M) {  // isNextAfterNumMatched(ci, 2): never true here
          int mLen = c_len(cg.op[ci + 2]);
          int tn = w.rp + multoffp, ts = w.start + w.clen;
          int tnm;
          int off = lookahead_offset(P, rd, ref, ts, tn, mLen, 2, &tnm);
          w.offset = off;
          nmoff += tnm;
          if (off != 0) {
            for (int k = 0; k < off; ++k) { app.push(rd.base(tn + k)); qsum += rd.q(tn + k); }
            qcnt += off;
          }
        }
        ci += 1;
      } else if (P.local_realign && n_cigar > ci + 1 && c_op(cg.op[ci + 1]) == OP_M) {  // isNextMatched
        int mLen = c_len(cg.op[ci + 1]);
        int tnm;
        int off = lookahead_offset(P, rd, ref, w.start + w.clen, w.rp, mLen, 2, &tnm);
        w.offset = off;
        nmoff += tnm;
        if (off != 0) {
          for (int k = 0; k < off; ++k) { app.push(rd.base(w.rp + k)); qsum += rd.q(w.rp + k); }
          qcnt += off;
        }
      }
      if (w.offset > 0) {
        key.push('&');
        for (int k = 0; k < app.n; ++k) key.push(app.c[k]);
      }
      // :1697-1713
      if (w.rp + w.offset >= rdh.l_seq) {
        qsum += q_before;
        qcnt += 1;
      } else {
        int q_after = rd.q(w.rp + w.offset);
        qsum += q_before > q_after ? q_before : q_after;
        qcnt += 1;
      }
      if (w.start >= R.start && w.start <= R.end) {
        // addVariationForDeletion :1036-1082
        int tp = w.re < rlen - w.re ? w.re + 1 : rlen - w.re;
        emit_event(sink, w, region_idx, read_idx, RV_EV_NI, w.start, &key, RV_EVF_PDEL, dir, tp, qsum, qcnt, mapq,
                   nm - nmoff, 0, 0, 0);
        for (int k = 0; k < w.clen; ++k) sink.cov(w.start + k);
      }
      w.start += w.clen + w.offset + multoffs;
      w.rp += w.offset + multoffp;
      w.re += w.offset + multoffp;
      return false;
    }

    // ---- match part :675-959 ------------------------------------------------------------------------
    if (fast_shape && w.offset == 0 && c_operator == OP_M) {
      // can any base of this run start a multi-nucleotide key?  (two mismatches within vext+1 bases; a
      // missing reference base or a non-ACGTN read base also sends the read down the exact path)
      bool plain = true;
      int last_mm = -1000;
      for (int i = 0; plain_hint < 0 && i < w.clen && plain; ++i) {
        char b = rd.base(w.rp + i);
        if (b == 'N') continue;
        if (!is_atgc(b)) { plain = false; break; }
        char rc = ref.at(w.start + i);
        if (rc != b) {
          if (i - last_mm <= P.vext + 1) plain = false;
          last_mm = i;
        }
      }
      if (plain) {
        fast->m_start = w.start;
        fast->m_len = (uint16_t)w.clen;
        fast->rp0 = (uint16_t)w.rp;
        fast->data_off16 = rdh.data_off16;
        fast->n_cigar = rdh.n_cigar;
        fast->nm = (int16_t)nm;
        fast->l_seq = (uint16_t)rdh.l_seq;
        fast->mapq = (uint8_t)mapq;
        fast->dir = dir ? 1 : 0;
        fast->read_idx = read_idx;
        w.start += w.clen;
        w.rp += w.clen;
        w.re += w.clen;
        if (w.start > R.end) return true;
        return false;
      }
    }
    int loop_from = w.offset, loop_end = w.clen;  // the per-base loop below covers [loop_from, loop_end) of the op
    SegDesc suffix;             // a plain tail of the stretch, emitted once the loop has reached it
    bool have_suffix = false;
    if (trim_after == 0 && w.clen - w.offset > 0 && nm >= 0 && nm <= 127 && rlen < 65536) {
      const bool indel_follows = ci + 1 < n_cigar && is_id(c_op(cg.op[ci + 1]));
      const int len = w.clen - w.offset;
      SegDesc sd;
      const int kind = sink.scan_segment(P, rd, ref, w.start, w.rp, len, indel_follows, &sd);
      if (kind == 1) {
        sd.m_start = w.start;
        sd.len = len;
        sd.rp = w.rp;
        sd.re = w.re;
        sd.rlen = rlen;
        if (sink.segment(sd, dir, mapq, nm)) {
          w.start += sd.len;
          w.rp += sd.len;
          w.re += sd.len;
          if (w.start > R.end) return true;
          return false;
        }
      } else if (kind == 2) {
        // plain apart from [p_first - (vext + 1), p_last]: what lies before and after goes to the gather kernel, the
        // literal loop only sees the flagged bases and the vext + 1 before them (stretches under 16 bases are not worth
        // a descriptor)
        int kp = sd.p_first - (vext + 1), ks = sd.p_last + 1;
#if !defined(__CUDA_ARCH__)
        if (getenv("RV_SPLIT_DEBUG")) fprintf(stderr, "split: start %d rp %d len %d flagged [%d,%d] n_mm %d\n", w.start, w.rp, len, sd.p_first, sd.p_last, sd.n_mm);
#endif
        if (kp < 16) kp = 0;
        if (len - ks < 16) ks = len;
        if (ks < len) {
          sub_mismatch_list(sd, ks, len, &suffix);
          suffix.m_start = w.start + ks;
          suffix.len = len - ks;
          suffix.rp = w.rp + ks;
          suffix.re = w.re + ks;
          suffix.rlen = rlen;
        }
        if (kp > 0) {
          SegDesc pre;
          sub_mismatch_list(sd, 0, kp, &pre);
          pre.m_start = w.start;
          pre.len = kp;
          pre.rp = w.rp;
          pre.re = w.re;
          pre.rlen = rlen;
          if (sink.segment(pre, dir, mapq, nm)) {
            w.start += kp;
            w.rp += kp;
            w.re += kp;
            loop_from += kp;  // (w.offset itself keeps its value: the op after a folded indel starts from it)
          }
        }
        if (ks < len) {  // (emitted after the loop: its bases record nm minus what the loop folded into keys, quirk A-8)
          have_suffix = true;
          loop_end = w.clen - (len - ks);
        }
      }
    }
    int nmoff = 0, moffset = 0;
    // The base, its quality and the reference base of the NEXT iteration are loaded one iteration ahead (read and
    // reference bytes never change during a pileup), so the common step "base equals the reference" waits on no load.
    int pf_rp = -1, pf_start = 0, pf_q = 0;
    char pf_base = 0, pf_ref = 0;
    WordCache c_seq, c_qual, c_ref;
    c_seq.tag = c_qual.tag = c_ref.tag = 0;
    c_seq.w = c_qual.w = c_ref.w = 0;
    for (;;) {
    for (int i = loop_from; i < loop_end; i++) {
      bool trim = false;
      if (trim_after != 0) trim = !dir ? (w.rp > trim_after) : (tlen - w.rp > trim_after);
      char ch1_, rc_;  // rc_: reference base at w.start, 0 outside the loaded window
      int q0_;
      if (pf_rp == w.rp && pf_start == w.start) {
        ch1_ = pf_base; q0_ = pf_q; rc_ = pf_ref;
      } else {
        ch1_ = rd.base(w.rp); q0_ = rd.q(w.rp); rc_ = ref.at(w.start);
      }
      pf_rp = w.rp + 1; pf_start = w.start + 1;
      pf_base = 0; pf_q = 0; pf_ref = 0;
      if (pf_rp < rd.lseq) {  // (pf_rp >= 1)
        const int b = cached_byte(rd.seq4 + (pf_rp >> 1), c_seq);
        pf_base = nt16_char((pf_rp & 1) ? (b & 15) : (b >> 4));
        pf_q = cached_byte(rd.qual + pf_rp, c_qual);
      }
      if (ref.has(pf_start)) pf_ref = (char)cached_byte((const uint8_t*)ref.bases + (pf_start - ref.base_pos), c_ref);
      const char ch1 = ch1_;
      if (ch1 == 'N') {
        w.start++;
        w.rp++;
        w.re++;
        continue;
      }
      if (rc_ == (char)0 || rc_ == ch1) {
        // !isHasAndNotEquals(ref, start, base): no multi-base key can start here (:711) and the indel-adjacent
        // forms need a mismatch (:779, :847) — the plain single-base observation of :884-937
        if (!trim && w.start >= r_start && w.start <= r_end) {
          const int tp = w.re < rlen - w.re ? w.re + 1 : rlen - w.re;
          const int al = allele_of(ch1);
          if (al >= 0) sink.single(w.start, al, dir, tp, q0_, mapq, nm - nmoff);
          else sink.unsupported();
          sink.cov(w.start);
        }
        w.start++;
        w.rp++;
        w.re++;
        continue;
      }
      int q = q0_;  // running SUM of qualities (a double in the reference; integer-valued)
      int qbases = 1, qibases = 0;
      Key s;
      s.clear();
      s.push(ch1);
      Key ss;
      ss.clear();
      bool start_with_deletion = false;
      char cur = ch1;  // rd.base(w.rp), carried along instead of re-loaded
      while ((w.start + 1) >= r_start && (w.start + 1) <= r_end && (i + 1) < w.clen && q >= goodq_i &&
             has_ne(ref, w.start, cur) && ref.at(w.start) != 'N') {
        if (rd.q(w.rp + 1) < goodq5_i) break;
        char nuc = rd.base(w.rp + 1);
        if (nuc == 'N') break;
        if (has_eq(ref, w.start + 1, 'N')) break;
        if (ref.at(w.start + 1) != nuc) {
          ss.push(nuc);
          q += rd.q(w.rp + 1);
          qbases++;
          w.rp++;
          w.re++;
          i++;
          w.start++;
          nmoff++;
          cur = nuc;
        } else {
          int ssn = 0;
          for (int ssi = 1; ssi <= vext; ssi++) {
            if (i + 1 + ssi >= w.clen) break;
            if (w.rp + 1 + ssi < rd.lseq && has_ne(ref, w.start + 1 + ssi, rd.base(w.rp + 1 + ssi))) {
              ssn = ssi + 1;
              break;
            }
          }
          if (ssn == 0) break;
          if (rd.q(w.rp + ssn) < goodq5_i) break;
          for (int ssi = 1; ssi <= ssn; ssi++) {
            ss.push(rd.base(w.rp + ssi));
            q += rd.q(w.rp + ssi);
            qbases++;
          }
          w.rp += ssn;
          w.re += ssn;
          i += ssn;
          w.start += ssn;
          cur = rd.base(w.rp);
        }
      }
      if (ss.n > 0) {
        s.push('&');
        for (int k = 0; k < ss.n; ++k) s.push(ss.c[k]);
      }
      int ddlen = 0;
      bool near_end = w.clen - i <= vext && P.local_realign && ci + 1 < n_cigar && ref.has(w.start) &&
                      (ss.n > 0 || cur != ref.at(w.start)) && rd.q(w.rp) >= goodq_i;
      if (near_end && c_op(cg.op[ci + 1]) == OP_D) {
        // :779-846
        while (i + 1 < w.clen) {
          s.push(rd.base(w.rp + 1));
          q += rd.q(w.rp + 1);
          qbases++;
          i++;
          w.rp++;
          w.re++;
          w.start++;
        }
        s.erase_first('&');
        Key pre;
        pre.clear();
        pre.push('-');
        pre.push_int(c_len(cg.op[ci + 1]));
        pre.push('&');
        s.prepend(pre.c, pre.n);
        start_with_deletion = true;
        ddlen = c_len(cg.op[ci + 1]);
        ci += 1;
        if (n_cigar > ci + 1 && c_op(cg.op[ci + 1]) == OP_I) {
          int next_len = c_len(cg.op[ci + 1]);
          s.push('^');
          for (int k = 0; k < next_len; ++k) s.push(rd.base(w.rp + 1 + k));
          for (int qi = 1; qi <= next_len; qi++) {
            q += rd.q(w.rp + 1 + qi);  // literal: starts one base late (:819-824)
            qibases++;
          }
          w.rp += next_len;
          w.re += next_len;
          ci += 1;
        }
        // isNextAfterNumMatched(ci, 1) reads the RECORD's cigar (same memory) at ci+1
        if (n_cigar > ci + 1 && c_op(cg.op[ci + 1]) == OP_M) {
          int tnm;
          int rpos = w.start + ddlen + 1, qpos = w.rp + 1;
          int off = lookahead_offset(P, rd, ref, rpos, qpos, c_len(cg.op[ci + 1]), 0, &tnm);
          if (off > 0) for (int k = 0; k < off; ++k) sink.cov(rpos + k);
          if (off != 0) {
            moffset = off;
            nmoff += tnm;
            s.push('&');
            for (int k = 0; k < off; ++k) s.push(rd.base(qpos + k));
            // the offset bases' qualities are never added (qualitySequence stays empty, A-7)
          }
        }
      } else if (near_end && c_op(cg.op[ci + 1]) == OP_I) {
        // :847-882
        while (i + 1 < w.clen) {
          s.push(rd.base(w.rp + 1));
          q += rd.q(w.rp + 1);
          qbases++;
          i++;
          w.rp++;
          w.re++;
          w.start++;
        }
        s.erase_first('&');
        int next_len = c_len(cg.op[ci + 1]);
        for (int k = 0; k < next_len; ++k) s.push(rd.base(w.rp + 1 + k));
        s.insert_at(next_len, '&');
        s.prepend("+", 1);
        for (int qi = 1; qi <= next_len; qi++) {
          q += rd.q(w.rp + 1 + qi);
          qibases++;
        }
        w.rp += next_len;
        w.re += next_len;
        ci += 1;
        qibases--;
        qbases++;
      }
      if (!trim) {
        const int pos = w.start - qbases + 1;
        if (pos >= r_start && pos <= r_end) {
          int tp = w.re < rlen - w.re ? w.re + 1 : rlen - w.re;
          if (s.n == 1) {
            int al = allele_of(ch1);
            if (al >= 0) sink.single(pos, al, dir, tp, q, mapq, nm - nmoff);
            else sink.unsupported();
          } else {
            int fl = (key_is_mnp(s) ? RV_EVF_MNP : 0) | (start_with_deletion ? RV_EVF_PDEL : 0);
            emit_event(sink, w, region_idx, read_idx, RV_EV_NI, pos, &s, fl, dir, tp, q, qbases + qibases, mapq,
                       nm - nmoff, 0, 0, 0);
          }
          for (int qi = 1; qi <= qbases; qi++) sink.cov(w.start - qi + 1);
          if (start_with_deletion)
            for (int qi = 1; qi < ddlen; qi++) sink.cov(w.start + qi);
        }
      }
      if (start_with_deletion) w.start += ddlen;
      w.start++;
      w.rp++;
      w.re++;
    }
    if (!have_suffix) break;
    // the loop stopped at the suffix' first base (no multi-nucleotide key reaches into it): hand the rest over, or walk
    // it literally after all when the sink has no room for another descriptor
    have_suffix = false;
    if (nm - nmoff >= 0 && sink.segment(suffix, dir, mapq, nm - nmoff)) {
      w.start += suffix.len;
      w.rp += suffix.len;
      w.re += suffix.len;
      break;
    }
    loop_from = loop_end;
    loop_end = w.clen;
    }
    if (moffset != 0) {
      w.offset = moffset;
      w.rp += moffset;
      w.start += moffset;
      w.re += moffset;
    }
    if (w.start > R.end) return true;
    return false;
  };
  bool active = has_work && n_cigar > 0;
  while (RV_WARP_ANY(active)) {
    if (active) {
      if (op_step()) active = false;
      else if (++ci >= n_cigar) active = false;
    }
  }
}

template <class Sink>
RV_HDN void process_read(const rv_params& P, const rv_region& R, int region_idx, const rv_read& rdh,
                         const uint8_t* pool, const RefView& ref, uint32_t read_idx, Sink& sink,
                         FastDesc* fast = (FastDesc*)0) {
  Prep pr;
  prepare_read(P, R, rdh, pool, ref, sink, fast != (FastDesc*)0, pr);
  if (!pr.ok) return;
  walk_read(P, R, region_idx, rdh, pool, ref, read_idx, sink, pr, fast, -1);
}

}  // namespace rvk
