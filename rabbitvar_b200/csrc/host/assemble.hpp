// assemble.hpp — host side of ToVarsBuilder::collectReferenceVariants (reference
// src/ToVarsBuilder.cpp:730-1098): turns the device's numeric variant records into full Variant
// objects (alleles, genotype, start/end, flanks), applies the output-stage filter
// (SimpleMode::output, src/modes/simpleMode.cpp:144-208 with Variant::varType/isGoodVar/adjComplex,
// include/Variant.h:99-247) and formats TSV lines (print_output_variant_simple, simpleMode.cpp:66-142).
// Only string work happens here; every number comes from the scoring kernel.
#pragma once
#include "pileup_model.hpp"
#include "fmt.hpp"
#include "../kernels/rv_core.cuh"
#include <stdio.h>
#include <string>
#include <vector>

namespace rvhost {

struct VariantOut {
  std::string key;  // descriptionString
  int cnt, fwd, rev;
  std::string bias;
  double freq, pmean, qual, mapq, qratio, hifreq, extrafreq, msi, nm;
  bool pstd, qstd;
  int shift3, msint, hicnt, hicov;
  std::string leftseq, rightseq;
  int start, end, ref_rev, ref_fwd, tcov;
  std::string genotype, varallele, refallele, vartype;
  double pvalue, oddratio;
  VariantOut() : cnt(0), fwd(0), rev(0), bias("0"), freq(0), pmean(0), qual(0), mapq(0), qratio(0), hifreq(0),
                 extrafreq(0), msi(0), nm(0), pstd(false), qstd(false), shift3(0), msint(0), hicnt(0), hicov(0),
                 start(0), end(0), ref_rev(0), ref_fwd(0), tcov(0), pvalue(1), oddratio(0) {}
};

struct PositionVars {
  int pos;
  bool has_ref;
  VariantOut ref;
  std::vector<VariantOut> variants;
};

inline std::string join_ref(const rvk::RefView& ref, int a, int b) {  // joinRef, VariationUtils.h:336-342
  std::string s;
  for (int i = a; i <= b; ++i) s.push_back(ref.at(i));
  return s;
}
inline bool starts_with(const std::string& s, const char* p) { return s.compare(0, strlen(p), p) == 0; }
inline void replace_first(std::string& s, const std::string& from, const std::string& to) {
  size_t p = s.find(from);
  if (p != std::string::npos) s.replace(p, from.size(), to);
}
inline bool is_atgc_char(char c) { return c == 'A' || c == 'T' || c == 'G' || c == 'C'; }
// regex_search(s, ".*&([ATGC]+).*") -> group 1 (greedy prefix: the LAST '&' that is followed by an ATGC run)
inline bool amp_atgc(const std::string& s, std::string* g1) {
  for (size_t i = s.size(); i-- > 0;) {
    if (s[i] == '&' && i + 1 < s.size() && is_atgc_char(s[i + 1])) {
      size_t j = i + 1;
      while (j < s.size() && is_atgc_char(s[j])) ++j;
      *g1 = s.substr(i + 1, j - i - 1);
      return true;
    }
  }
  return false;
}
// regex_search(s, "#(.+)\\^(.+)")
inline bool hash_caret(const std::string& s, std::string* g1, std::string* g2) {
  size_t h = s.find('#');
  while (h != std::string::npos) {
    // greedy (.+): last '^' after h+1 with at least one char after it
    for (size_t c = s.size(); c-- > h + 2;) {
      if (s[c] == '^' && c + 1 < s.size()) {
        *g1 = s.substr(h + 1, c - h - 1);
        *g2 = s.substr(c + 1);
        return true;
      }
    }
    h = s.find('#', h + 1);
  }
  return false;
}
inline std::string key_of(const rv_variant& v, const std::vector<rv_patch_entry>& patch) {
  if (v.key_kind == 0) return std::string(1, "ACGT"[v.key_id]);
  const rv_patch_entry& e = patch[(size_t)v.key_id];
  return std::string(e.key, e.keylen);
}

inline void fill_numeric(VariantOut& o, const rv_variant& v, const std::string& key) {
  o.key = key;
  o.cnt = v.cnt; o.fwd = v.fwd; o.rev = v.rev;
  o.freq = v.freq; o.pmean = v.pmean; o.qual = v.qual; o.mapq = v.mapq; o.qratio = v.qratio; o.hifreq = v.hifreq;
  o.extrafreq = v.extrafreq; o.nm = v.nm; o.pstd = v.pstd; o.qstd = v.qstd; o.hicnt = v.hicnt; o.hicov = v.hicov;
  o.pvalue = v.pvalue; o.oddratio = v.oddratio;
}

// Records of ONE position, device order (variants by rank, reference record anywhere).
inline void assemble_position(const rv_params& P, const rv_variant* recs, int n, const std::vector<rv_patch_entry>& patch,
                              const rvk::RefView& ref, int chr_len, PositionVars* out) {
  const int position = recs[0].pos;
  out->pos = position;
  out->has_ref = false;
  out->variants.clear();
  const rv_variant* refrec = NULL;
  std::vector<const rv_variant*> vars;
  for (int i = 0; i < n; ++i) {
    if (recs[i].is_ref) refrec = &recs[i];
    else vars.push_back(&recs[i]);
  }
  if (refrec) {
    out->has_ref = true;
    fill_numeric(out->ref, *refrec, key_of(*refrec, patch));
    out->ref.bias = std::to_string((int)refrec->bias_ref);
  }
  std::string genotype1;
  if (refrec && out->ref.freq >= P.freq) genotype1 = out->ref.key;
  else if (!vars.empty()) genotype1 = key_of(*vars[0], patch);
  else if (refrec) genotype1 = out->ref.key;
  if (starts_with(genotype1, "+")) genotype1 = "+" + std::to_string(genotype1.size() - 1);
  if (vars.empty()) {
    if (refrec) {  // ToVarsBuilder.cpp:1066-1090
      VariantOut& r = out->ref;
      r.tcov = refrec->tcov;
      r.cnt = 0; r.freq = 0; r.ref_fwd = refrec->ref_fwd; r.ref_rev = refrec->ref_rev; r.fwd = 0; r.rev = 0;
      r.msi = 0; r.msint = 0; r.bias += ";0"; r.shift3 = 0; r.start = position; r.end = position;
      std::string rb = ref.has(position) ? std::string(1, ref.at(position)) : std::string();
      r.refallele = rb; r.varallele = rb; r.genotype = rb + "/" + rb;
    }
    return;
  }
  for (size_t vi = 0; vi < vars.size(); ++vi) {
    const rv_variant& rec = *vars[vi];
    VariantOut v;
    const std::string vn = key_of(rec, patch);
    fill_numeric(v, rec, vn);
    std::string genotype2 = vn;
    if (starts_with(genotype2, "+")) genotype2 = "+" + std::to_string(genotype2.size() - 1);
    int dellen = 0;
    if (vn.size() > 1 && vn[0] == '-' && vn[1] >= '0' && vn[1] <= '9')
      for (size_t i = 1; i < vn.size() && vn[i] >= '0' && vn[i] <= '9'; ++i) dellen = dellen * 10 + (vn[i] - '0');
    int ep = position;
    if (starts_with(vn, "-")) ep = position + dellen - 1;
    std::string refallele, varallele;
    int shift3 = rec.shift3;
    int sp = position;
    if (starts_with(vn, "+")) {
      if (P.move3) { sp += shift3; ep += shift3; }
      refallele = ref.has(position) ? std::string(1, ref.at(position)) : std::string();
      varallele = refallele + vn.substr(1);
    } else if (starts_with(vn, "-")) {
      if (dellen < 1000) {
        varallele = vn;
        size_t k = 1;
        while (k < varallele.size() && varallele[k] >= '0' && varallele[k] <= '9') ++k;
        if (k > 1) varallele.erase(0, k);  // "^-\d+"
      }
      if (vn.find('&') == std::string::npos && vn.find('#') == std::string::npos && vn.find('^') == std::string::npos) {
        if (P.move3) sp += shift3;
        if (varallele != "<DEL>") varallele = ref.has(position - 1) ? std::string(1, ref.at(position - 1)) : std::string();
        refallele = ref.has(position - 1) ? std::string(1, ref.at(position - 1)) : std::string();
        sp--;
      }
      if (dellen < 1000) refallele += join_ref(ref, position, position + dellen - 1);
    } else {
      refallele = ref.has(position) ? std::string(1, ref.at(position)) : std::string();
      varallele = vn;
    }
    std::string extra;
    if (amp_atgc(vn, &extra)) {
      replace_first(varallele, "&", "");
      std::string tch = join_ref(ref, ep + 1, ep + (int)extra.size());
      refallele += tch;
      genotype1 += tch;
      ep += (int)extra.size();
      std::string vextra;
      if (amp_atgc(varallele, &vextra)) {
        replace_first(varallele, "&", "");
        tch = join_ref(ref, ep + 1, ep + (int)vextra.size());
        refallele += tch;
        genotype1 += tch;
        ep += (int)vextra.size();
      }
      if (starts_with(vn, "+")) {
        refallele = refallele.substr(1);
        varallele = varallele.substr(1);
        sp++;
      }
    }
    std::string mseq, tail;
    if (hash_caret(vn, &mseq, &tail)) {
      ep += (int)mseq.size();
      refallele += join_ref(ref, ep - (int)mseq.size() + 1, ep);
      if (!tail.empty() && tail[0] >= '0' && tail[0] <= '9') {
        int d = atoi(tail.c_str());
        refallele += join_ref(ref, ep + 1, ep + d);
        ep += d;
      }
      replace_first(varallele, "#", "");
      size_t c = varallele.find('^');  // "\^(\d+)?" first match
      if (c != std::string::npos) {
        size_t e = c + 1;
        while (e < varallele.size() && varallele[e] >= '0' && varallele[e] <= '9') ++e;
        varallele.erase(c, e - c);
      }
      replace_first(genotype1, "#", "m");
      replace_first(genotype1, "^", "i");
      replace_first(genotype2, "#", "m");
      replace_first(genotype2, "^", "i");
    }
    {  // CARET_ATGNC "\^([ATGNC]+)"
      bool m = false;
      for (size_t i = 0; i + 1 < vn.size(); ++i)
        if (vn[i] == '^' && (is_atgc_char(vn[i + 1]) || vn[i + 1] == 'N')) m = true;
      if (m) {
        replace_first(varallele, "^", "");
        replace_first(genotype1, "^", "i");
        replace_first(genotype2, "^", "i");
      }
    }
    v.leftseq = join_ref(ref, sp - 20 < 1 ? 1 : sp - 20, sp - 1);
    v.rightseq = join_ref(ref, ep + 1, ep + 20 > chr_len ? chr_len : ep + 20);
    std::string genotype = genotype1 + "/" + genotype2;
    std::string g2;
    for (size_t i = 0; i < genotype.size(); ++i) {
      char ch = genotype[i];
      if (ch == '&' || ch == '#') continue;
      g2.push_back(ch == '^' ? 'i' : ch);
    }
    v.genotype = g2;
    v.msi = rec.msi;
    v.msint = rec.msint;
    v.shift3 = shift3;
    v.start = sp;
    v.end = ep;
    v.refallele = refallele;
    v.varallele = varallele;
    v.tcov = rec.tcov;
    v.ref_fwd = rec.ref_fwd;
    v.ref_rev = rec.ref_rev;
    v.bias = (refrec ? std::to_string((int)rec.bias_ref) : std::string("0")) + ";" + std::to_string((int)rec.bias_var);
    out->variants.push_back(v);
  }
}

// Variant::varType, include/Variant.h:167-188
inline std::string var_type(const VariantOut& v) {
  const std::string &r = v.refallele, &a = v.varallele;
  if (r.size() == 1 && a.size() == 1) return "SNV";
  if (a.size() == 5 && a[0] == '<' && a[4] == '>') return a.substr(1, 3);
  if (r.empty() || a.empty()) return "Complex";
  if (r[0] != a[0]) return "Complex";
  if (r.size() == 1 && a.size() > 1 && a.compare(0, r.size(), r) == 0) return "Insertion";
  if (r.size() > 1 && a.size() == 1 && r.compare(0, a.size(), a) == 0) return "Deletion";
  return "Complex";
}

// Variant::isGoodVar, include/Variant.h:197-247 (splice set is empty: no N CIGAR ops reach this path)
inline bool is_good_var(const VariantOut& v, const VariantOut* refv, const std::string& type, const rv_params& P) {
  if (v.refallele.empty()) return false;
  if (v.freq < P.freq || v.hicnt < P.minr || v.pmean < P.read_pos_filter || v.qual < P.goodq) return false;
  if (refv != NULL && refv->hicnt > P.minr && v.freq < 0.25) {
    double d = v.mapq + v.refallele.size() + v.varallele.size();
    double f = (1 + d) / (refv->mapq + 1);
    if ((d - 2 < 5 && refv->mapq > 20) || f < 0.25) return false;
  }
  if (v.qratio < P.qratio) return false;
  if (v.freq > 0.30) return true;
  if (v.mapq < P.mapq) return false;
  if (v.bias == "2;1" && v.freq < 0.20) {
    if (type.empty() || type == "SNV" || (v.refallele.size() < 3 && v.varallele.size() < 3)) return false;
  }
  return true;
}

inline std::string sub_from(const std::string& s, int idx) {  // vc_substr(str, idx)
  if (idx >= 0) return (size_t)idx <= s.size() ? s.substr(idx) : std::string();
  if ((int)s.size() + idx < 0) return "";
  return s.substr(s.size() + idx);
}
inline std::string sub_len(const std::string& s, int begin, int len) {  // vc_substr(str, begin, len)
  if (begin < 0) begin = (int)s.size() + begin;
  if (begin < 0 || (size_t)begin > s.size()) return "";
  if (len > 0) return s.substr(begin, len);
  if (len == 0) return "";
  len = (int)s.size() + len - begin;
  if (len < 0) return "";
  return s.substr(begin, len);
}

// Variant::adjComplex, include/Variant.h:99-132
inline void adj_complex(VariantOut& v) {
  std::string refAllele = v.refallele, varAllele = v.varallele;
  if (!varAllele.empty() && varAllele[0] == '<') return;
  size_t n = 0;
  while (refAllele.size() - n > 1 && varAllele.size() - n > 1 && refAllele[n] == varAllele[n]) n++;
  if (n > 0) {
    v.start += (int)n;
    v.refallele = sub_from(refAllele, (int)n);
    v.varallele = sub_from(varAllele, (int)n);
    v.leftseq += sub_len(refAllele, 0, (int)n);
    v.leftseq = sub_from(v.leftseq, (int)n);
  }
  refAllele = v.refallele;
  varAllele = v.varallele;
  n = 1;
  while (refAllele.size() - n > 0 && varAllele.size() - n > 0 &&
         sub_len(refAllele, -(int)n, 1) == sub_len(varAllele, -(int)n, 1))
    n++;
  if (n > 1) {
    v.end -= (int)n - 1;
    v.refallele = sub_len(refAllele, 0, 1 - (int)n);
    v.varallele = sub_len(varAllele, 0, 1 - (int)n);
    v.rightseq = sub_len(refAllele, 1 - (int)n, (int)n - 1) + sub_len(v.rightseq, 0, 1 - (int)n);
  }
}

// print_output_variant_simple, simpleMode.cpp:66-142
inline std::string format_simple(const VariantOut& v, const std::string& sample, const std::string& gene,
                                 const std::string& chr, int rstart, int rend, bool fisher) {
  std::string s;
  auto add = [&](const std::string& f) { s += f; s += '\t'; };
  add(sample); add(gene); add(chr);
  add(std::to_string(v.start)); add(std::to_string(v.end)); add(v.refallele); add(v.varallele);
  add(std::to_string(v.tcov)); add(std::to_string(v.cnt)); add(std::to_string(v.ref_fwd)); add(std::to_string(v.ref_rev));
  add(std::to_string(v.fwd)); add(std::to_string(v.rev)); add(v.genotype.empty() ? "0" : v.genotype);
  add(f6(v.freq)); add(v.bias); add(f6(v.pmean)); add(v.pstd ? "1" : "0");
  add(f6(v.qual)); add(v.qstd ? "1" : "0");
  if (fisher) { add(f6(v.pvalue)); add(f6(v.oddratio)); }
  add(f6(v.mapq)); add(f6(v.qratio)); add(f6(v.hifreq));
  add(f6(v.extrafreq)); add(std::to_string(v.shift3)); add(f6(v.msi));
  add(std::to_string(v.msint)); add(v.nm > 0 ? f6(v.nm) : std::to_string(0)); add(std::to_string(v.hicnt));
  add(std::to_string(v.hicov)); add(v.leftseq.empty() ? "0" : v.leftseq); add(v.rightseq.empty() ? "0" : v.rightseq);
  add(chr + ":" + std::to_string(rstart) + "-" + std::to_string(rend)); add(v.vartype);
  add(f6(0.0));  // duprate: CigarParser::process forces 0 (parseCigar.cpp:432)
  s += "0\n";                // sv placeholder
  return s;
}

// SimpleMode::output for one position (simpleMode.cpp:144-208); appends TSV lines.
inline void output_position_simple(const rv_params& P, PositionVars& pv, const std::string& sample,
                                   const std::string& gene, const std::string& chr, int rstart, int rend,
                                   std::string* out) {
  if (pv.pos < rstart || pv.pos > rend) return;
  std::vector<VariantOut*> vrefs;
  if (pv.variants.empty()) {
    if (!P.pileup) return;
    if (!pv.has_ref) return;
    pv.ref.vartype = "";
    vrefs.push_back(&pv.ref);
  } else {
    for (size_t i = 0; i < pv.variants.size(); ++i) {
      VariantOut& v = pv.variants[i];
      if (v.refallele.find('N') != std::string::npos) continue;
      v.vartype = var_type(v);
      if (!is_good_var(v, pv.has_ref ? &pv.ref : NULL, v.vartype, P))
        if (!P.pileup) continue;
      vrefs.push_back(&v);
    }
  }
  for (size_t i = 0; i < vrefs.size(); ++i) {
    VariantOut& v = *vrefs[i];
    if (v.vartype == "Complex") adj_complex(v);
    *out += format_simple(v, sample, gene, chr, rstart, rend, P.fisher != 0);
  }
}

inline void dump_variant_line(FILE* f, const char* tag, int pos, const VariantOut& v) {
  fprintf(f, "%s\t%d\t%s\t%d\t%d\t%d\t%s\t%.17g\t%.17g\t%d\t%.17g\t%d\t%.17g\t%.17g\t%.17g\t%.17g\t%d\t%.17g\t%d\t%.17g\t%d\t%d\t%s\t%s\t%d\t%d\t%d\t%d\t%d\t%s\t%s\t%s\n",
          tag, pos, v.key.c_str(), v.cnt, v.fwd, v.rev, v.bias.c_str(), v.freq, v.pmean, v.pstd ? 1 : 0, v.qual,
          v.qstd ? 1 : 0, v.mapq, v.qratio, v.hifreq, v.extrafreq, v.shift3, v.msi, v.msint, v.nm, v.hicnt, v.hicov,
          v.leftseq.empty() ? "." : v.leftseq.c_str(), v.rightseq.empty() ? "." : v.rightseq.c_str(), v.start, v.end,
          v.ref_rev, v.ref_fwd, v.tcov, v.genotype.empty() ? "." : v.genotype.c_str(),
          v.varallele.empty() ? "." : v.varallele.c_str(), v.refallele.empty() ? "." : v.refallele.c_str());
}

// Writes the V.REF / V.VAR lines of oracle/ref_dump.cpp for a sorted variant list (all regions).
inline void dump_variants(FILE* out, const rv_params& P, const std::vector<rv_variant>& v,
                          const std::vector<std::vector<rv_patch_entry> >& patches, const std::vector<rv_region>& regs,
                          const rvk::RefView& ref_in, const std::string& chr) {
  (void)chr;
  rvk::RefView ref = ref_in;
  for (size_t i = 0; i < v.size();) {
    size_t j = i;
    while (j < v.size() && v[j].region == v[i].region && v[j].pos == v[i].pos) ++j;
    const rv_region& R = regs[(size_t)v[i].region];
    ref.lo = R.ref_lo;
    ref.hi = R.ref_hi;
    PositionVars pv;
    assemble_position(P, &v[i], (int)(j - i), patches[(size_t)v[i].region], ref, R.chr_len, &pv);
    if (pv.has_ref) dump_variant_line(out, "V.REF", pv.pos, pv.ref);
    for (size_t k = 0; k < pv.variants.size(); ++k) dump_variant_line(out, "V.VAR", pv.pos, pv.variants[k]);
    i = j;
  }
}

}  // namespace rvhost
