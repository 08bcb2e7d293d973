// somatic.hpp — host side of the paired (tumor | normal) mode: the per-position join of the two samples'
// variant lists, the classification labels and the 55/63-column TSV line
// (reference src/modes/somaticMode.cpp:151-309 print_output_variant_simple, :311-352 output, :360-386
// callingForOneSample / callingForBothSamples, :393-525 printVariationsFromFirstSample, :534-597
// printVariationsFromSecondSample, :607-631 determinateType; Variant::isNoise include/Variant.h:75-94).
// combineAnalysis (:643-...) only ever runs for structural-variant keys (descriptions containing '<'), which
// this path never produces, so it has no counterpart here.
#pragma once
#include "assemble.hpp"
#include "../kernels/rv_score.cuh"
#include <math.h>

namespace rvhost {

// Variant::isNoise, include/Variant.h:75-94 (zeroes the variant's counts when it is noise)
inline bool is_noise(VariantOut& v, const rv_params& P) {
  const double qual = v.qual;
  if (((qual < 4.5 || (qual < 12 && !v.qstd)) && v.cnt <= 3) || (qual < P.goodq && v.freq < 2 * P.lofreq && v.cnt <= 1)) {
    v.tcov -= v.cnt;
    v.cnt = 0;
    v.fwd = 0;
    v.rev = 0;
    v.freq = 0;
    v.hifreq = 0;
    return true;
  }
  return false;
}

// put_fisher_ext_and_odds, somaticMode.cpp:130-149 (single-precision odds, as written)
inline void put_fisher_and_odds(std::string& s, int ref_fwd, int ref_rev, int alt_fwd, int alt_rev) {
  // lgamma(n + 1) table for the host-side Fisher tests (three per printed line)
  static const std::vector<double> table = [] {
    std::vector<double> t(1 << 16);
    for (size_t i = 0; i < t.size(); ++i) t[i] = lgamma((double)i + 1.0);
    return t;
  }();
  rvk::LgTable lg;
  lg.t = table.data();
  lg.n = (int)table.size();
  double l, r, two;
  rvk::fisher_exact(lg, ref_fwd, ref_rev, alt_fwd, alt_rev, &l, &r, &two);
  append_f6(s, two);
  s += '\t';
  const float t_ref_fwd = ref_fwd + 0.5, t_ref_rev = ref_rev + 0.5, t_alt_fwd = alt_fwd + 0.5, t_alt_rev = alt_rev + 0.5;
  const float ad = t_ref_fwd * t_alt_rev;
  const float bc = t_ref_rev * t_alt_fwd;
  const float ratio = std::log(ad / bc + bc / ad);
  append_f6(s, (double)ratio);
  s += '\t';
}

inline void sample_block(std::string& s, const VariantOut* v, const VariantOut* tumor_for_genotype) {
  if (!v) {
    s += "0\t0\t0\t0\t0\t0\t0\t0\t0\t0\t0\t0\t0\t0\t0\t0\t0\t0\t";
    return;
  }
  auto add = [&](const std::string& f) { s += f; s += '\t'; };
  add(std::to_string(v->tcov)); add(std::to_string(v->cnt)); add(std::to_string(v->ref_fwd)); add(std::to_string(v->ref_rev));
  add(std::to_string(v->fwd)); add(std::to_string(v->rev));
  // the normal block prints the TUMOR genotype when its own is set (somaticMode.cpp:205, as written)
  add(v->genotype.empty() ? "0" : (tumor_for_genotype ? tumor_for_genotype->genotype : v->genotype));
  add(f6(v->freq)); add(v->bias); add(f6(v->pmean)); add(v->pstd ? "1" : "0");
  add(f6(v->qual)); add(v->qstd ? "1" : "0"); add(f6(v->mapq)); add(f6(v->qratio));
  add(f6(v->hifreq)); add(f6(v->extrafreq)); add(f6(v->nm));
}

// print_output_variant_simple, somaticMode.cpp:151-309
inline std::string format_somatic(const VariantOut* begin, const VariantOut* end, const VariantOut* tumor,
                                  const VariantOut* normal, const std::string& sample, const std::string& gene,
                                  const std::string& chr, int rstart, int rend, const std::string& label, bool fisher) {
  std::string s;
  auto add = [&](const std::string& f) { s += f; s += '\t'; };
  add(sample); add(gene); add(chr);
  if (begin) { add(std::to_string(begin->start)); add(std::to_string(begin->end)); add(begin->refallele); add(begin->varallele); }
  else s += "0\t0\t0\t0\t";
  sample_block(s, tumor, NULL);
  if (fisher) {
    if (tumor) put_fisher_and_odds(s, tumor->ref_fwd, tumor->ref_rev, tumor->fwd, tumor->rev);
    else s += "0\t0\t";
  }
  if (normal && !normal->genotype.empty() && !tumor) {
    // the reference dereferences a null tumor variant here; no call site reaches it with real data
    sample_block(s, normal, normal);
  } else {
    sample_block(s, normal, tumor);
  }
  if (fisher) {
    if (normal) put_fisher_and_odds(s, normal->ref_fwd, normal->ref_rev, normal->fwd, normal->rev);
    else s += "0\t0\t";
  }
  if (end) {
    add(std::to_string(end->shift3)); add(f6(end->msi)); add(std::to_string(end->msint));
    add(end->leftseq.empty() ? "0" : end->leftseq); add(end->rightseq.empty() ? "0" : end->rightseq);
  } else {
    s += "\t\t\t\t\t";
  }
  add(chr + ":" + std::to_string(rstart) + "-" + std::to_string(rend));
  add(label);
  if (begin) add(begin->vartype);
  else s += "\t";
  add(tumor ? f6(0.0) : "0");  // duprate is forced to 0 by CigarParser::process (parseCigar.cpp:432)
  add("0");
  add(normal ? f6(0.0) : "0");
  add("0");
  if (fisher) {
    const int v1t = tumor ? tumor->tcov : 0, v1v = tumor ? tumor->cnt : 0, v2t = normal ? normal->tcov : 0, v2v = normal ? normal->cnt : 0;
    int tref = v1t - v1v, rref = v2t - v2v;
    if (tref < 0) tref = 0;
    if (rref < 0) rref = 0;
    put_fisher_and_odds(s, v1v, tref, v2v, rref);
    const double tumor_vaf = tumor ? tumor->freq : 0, normal_vaf = normal ? normal->freq : 0;
    const double lo = std::log(std::max(tumor_vaf, 0.0001) / std::max(normal_vaf, 0.0001));
    add(f6(lo));
    add(f6(std::log((static_cast<float>(v1v) + 0.5) / (static_cast<float>(v2v) + 0.5))));
  }
  s += "\n";
  return s;
}

inline VariantOut* find_by_key(PositionVars& pv, const std::string& key) {  // varDescriptionStringToVariants
  // later variants with the same description string overwrite earlier ones in the reference's map
  VariantOut* hit = NULL;
  for (size_t i = 0; i < pv.variants.size(); ++i)
    if (pv.variants[i].key == key) hit = &pv.variants[i];
  return hit;
}
inline bool minus_num_num(const std::string& s) {  // regex_search(s, "-\\d\\d")
  for (size_t i = 0; i + 2 < s.size(); ++i)
    if (s[i] == '-' && s[i + 1] >= '0' && s[i + 1] <= '9' && s[i + 2] >= '0' && s[i + 2] <= '9') return true;
  return false;
}

// determinateType, somaticMode.cpp:607-631
inline std::string determinate_type(PositionVars& v2, const VariantOut& standard, VariantOut& cmp, const rv_params& P) {
  std::string type;
  if (is_good_var(cmp, v2.has_ref ? &v2.ref : NULL, standard.vartype, P)) {
    if (standard.freq > (1 - P.lofreq) && cmp.freq < 0.8 && cmp.freq > 0.2) type = "LikelyLOH";
    else if (cmp.freq < P.lofreq || cmp.cnt <= 1) type = "LikelySomatic";
    else type = "Germline";
  } else {
    if (cmp.freq < P.lofreq || cmp.cnt <= 1) type = "LikelySomatic";
    else type = "AFDiff";
  }
  if (is_noise(cmp, P) && standard.vartype == "SNV") type = "StrongSomatic";
  return type;
}

struct SomaticCtx {
  const rv_params& P;
  const std::string &sample, &gene, &chr;
  int rstart, rend;
  std::string* out;
  void print(const VariantOut* b, const VariantOut* e, const VariantOut* t, const VariantOut* n, const std::string& label) {
    *out += format_somatic(b, e, t, n, sample, gene, chr, rstart, rend, label, P.fisher != 0);
  }
};

// callingForOneSample, somaticMode.cpp:360-386
inline void calling_for_one_sample(SomaticCtx& C, PositionVars& v, bool is_first_cover, const std::string& label) {
  for (size_t i = 0; i < v.variants.size(); ++i) {
    VariantOut& var = v.variants[i];
    var.vartype = var_type(var);
    if (!is_good_var(var, v.has_ref ? &v.ref : NULL, var.vartype, C.P)) continue;
    if (var.vartype == "Complex") adj_complex(var);
    if (is_first_cover) C.print(&var, &var, NULL, &var, label);
    else C.print(&var, &var, &var, NULL, label);
  }
}

// printVariationsFromFirstSample, somaticMode.cpp:393-525
inline void print_from_first(SomaticCtx& C, PositionVars& v1, PositionVars& v2) {
  const rv_params& P = C.P;
  size_t n = 0;
  // the loop test reads `vartype` before it is assigned: it is still empty there (somaticMode.cpp:396-398)
  while (n < v1.variants.size() && is_good_var(v1.variants[n], v1.has_ref ? &v1.ref : NULL, v1.variants[n].vartype, P)) {
    VariantOut& vref = v1.variants[n];
    const std::string nt = vref.key;
    vref.vartype = var_type(vref);
    if (vref.vartype == "Complex") adj_complex(vref);
    VariantOut* v2nt = find_by_key(v2, nt);
    if (v2nt) {
      const std::string type = determinate_type(v2, vref, *v2nt, P);
      C.print(&vref, v2nt, &vref, v2nt, type);
    } else {  // sample 1 only, should be strong somatic
      VariantOut blank;
      const VariantOut* for_print = NULL;
      if (!v2.variants.empty()) {
        blank.tcov = v2.variants[0].tcov;
        blank.ref_fwd = v2.variants[0].ref_fwd;
        blank.ref_rev = v2.variants[0].ref_rev;
        for_print = &blank;
      } else if (v2.has_ref) {
        for_print = &v2.ref;
      }
      // (combineAnalysis is only reached for keys containing '<': never on this path)
      C.print(&vref, &vref, &vref, for_print, "StrongSomatic");
    }
    n++;
  }
  if (n == 0) {
    if (v2.variants.empty()) return;
    for (size_t i = 0; i < v2.variants.size(); ++i) {
      VariantOut& v2var = v2.variants[i];
      v2var.vartype = var_type(v2var);
      if (!is_good_var(v2var, v2.has_ref ? &v2.ref : NULL, v2var.vartype, P)) continue;
      const std::string nt = v2var.key;  // potential LOH
      VariantOut* v1nt = find_by_key(v1, nt);
      if (v1nt) {
        const std::string type = v1nt->freq < P.lofreq ? "LikelyLOH" : "Germline";
        if (v2var.vartype == "Complex") adj_complex(*v1nt);
        v1nt->vartype = var_type(*v1nt);
        C.print(v1nt, &v2var, v1nt, &v2var, type);
      } else {
        const VariantOut* v1var = v1.variants.empty() ? NULL : &v1.variants[0];
        const int tcov = v1var && v1var->tcov != 0 ? v1var->tcov : 0;
        const int fwd = v1.has_ref ? v1.ref.fwd : 0, rev = v1.has_ref ? v1.ref.rev : 0;
        const std::string genotype = v1var ? v1var->genotype : (v1.has_ref ? v1.ref.key + "/" + v1.ref.key : "N/N");
        if (v2var.vartype == "Complex") adj_complex(v2var);
        VariantOut for_print;
        for_print.tcov = tcov;
        for_print.ref_fwd = fwd;
        for_print.ref_rev = rev;
        for_print.genotype = genotype;
        C.print(&v2var, &v2var, &for_print, &v2var, "StrongLOH");
      }
    }
  }
}

// printVariationsFromSecondSample, somaticMode.cpp:534-597
inline void print_from_second(SomaticCtx& C, PositionVars& v1, PositionVars& v2) {
  for (size_t i = 0; i < v2.variants.size(); ++i) {
    VariantOut& v2var = v2.variants[i];
    v2var.vartype = var_type(v2var);
    if (!is_good_var(v2var, v2.has_ref ? &v2.ref : NULL, v2var.vartype, C.P)) continue;
    VariantOut* v1nt = find_by_key(v1, v2var.key);  // potential LOH
    if (v1nt) v1nt->cnt = 0;
    const VariantOut* for_print = v1.has_ref ? &v1.ref : NULL;
    if (v2var.vartype == "Complex") adj_complex(v2var);
    C.print(&v2var, &v2var, for_print, &v2var, "StrongLOH");
  }
}

// SomaticMode::output for one position of the tumor sample (somaticMode.cpp:311-352): `v2` is NULL when the
// normal sample has nothing at the position.
inline void output_position_somatic(const rv_params& P, PositionVars& v1, PositionVars* v2, const std::string& sample,
                                    const std::string& gene, const std::string& chr, int rstart, int rend, std::string* out) {
  if (v1.pos < rstart || v1.pos > rend) return;
  SomaticCtx C = {P, sample, gene, chr, rstart, rend, out};
  if (!v2) {
    calling_for_one_sample(C, v1, false, "SampleSpecific");
    return;
  }
  if (v1.variants.empty() && v2->variants.empty()) return;
  if (!v1.variants.empty()) print_from_first(C, v1, *v2);
  else print_from_second(C, v1, *v2);
}

}  // namespace rvhost
