// ref_dump — drives the UNMODIFIED reference's per-region pipeline (the objects built from
// /root/reference/src by oracle/Makefile) and dumps its intermediate tables so the CUDA path can be
// compared stage by stage.  TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline leg may execute this.
//
// It repeats the call sequence of SimpleMode::one_region_run (reference src/modes/simpleMode.cpp:18-64)
// / one_region_run_somt (src/modes/somaticMode.cpp:83-127) through the reference's public classes and
// writes, for every region:
//   stage C  (after CigarParser::process,        src/parseCigar.cpp:410)
//   stage R  (after VariationRealigner::process, src/VariationRealigner.cpp:135)
//   stage V  (after ToVarsBuilder::process,      src/ToVarsBuilder.cpp:97)
// as sorted text lines into  $RV_DUMP (default ref_dump.txt).  Command line = the reference's own
// (parsed by its cmdParse, src/Launcher.cpp:295).  Doubles are printed with %.17g (round-trip exact).
//
// Line formats (tab separated):
//   REGION  sample chr start end
//   {C|R}.NI pos key cnt fwd rev sumPos sumQual sumMapq sumNm lo hi pstd qstd extracnt      (nonInsertionVariants)
//   {C|R}.IN pos key ...same...                                                             (insertionVariants)
//   {C|R}.COV pos cov                                                                        (refCoverage)
//   {C|R}.SC5|SC3 pos cnt fwd rev sumPos sumQual sumMapq sumNm lo hi used                    (soft clips)
//   {C|R}.SCNT 5|3 pos idx base count  /  .SCSEQ 5|3 pos idx base cnt fwd rev sumPos sumQual sumMapq sumNm lo hi
//   C.PINS / C.PDEL / C.MNP pos key count ;  C.MAXRL n
//   V.REF / V.VAR pos key + all Variant fields
#include "Launcher.h"
#include "RegionBuilder.h"
#include "recordPreprocessor.h"
#include "parseCigar.h"
#include "VariationRealigner.h"
#include "ToVarsBuilder.h"
#include "patterns.h"
#include <assert.h>
#include <algorithm>
#include <vector>
#include <string>

Configuration* cmdParse(int argc, char* argv[]);  // reference src/Launcher.cpp:295

static FILE* OUT = NULL;
typedef robin_hood::unordered_map<int, VariationMap*> PosVarMap;

static void dump_variation_fields(const Variation* v) {
  fprintf(OUT, "%d\t%d\t%d\t%.17g\t%.17g\t%.17g\t%.17g\t%d\t%d\t%d\t%d\t%d", v->varsCount, v->varsCountOnForward,
          v->varsCountOnReverse, v->meanPosition, v->meanQuality, v->meanMappingQuality, v->numberOfMismatches,
          v->lowQualityReadsCount, v->highQualityReadsCount, v->pstd ? 1 : 0, v->qstd ? 1 : 0, v->extracnt);
}

static void dump_posvarmap(const char* tag, PosVarMap* m) {
  std::vector<std::pair<int, std::string> > keys;
  for (auto& pv : *m)
    for (auto& kv : pv.second->variation_map) keys.push_back(std::make_pair(pv.first, kv.first));
  std::sort(keys.begin(), keys.end());
  for (auto& k : keys) {
    Variation* v = m->at(k.first)->variation_map.at(k.second);
    fprintf(OUT, "%s\t%d\t%s\t", tag, k.first, k.second.c_str());
    dump_variation_fields(v);
    fputc('\n', OUT);
  }
}

static void dump_cov(const char* tag, robin_hood::unordered_map<int, int>* cov) {
  std::vector<std::pair<int, int> > v;
  for (auto& e : *cov) v.push_back(std::make_pair(e.first, e.second));
  std::sort(v.begin(), v.end());
  for (auto& p : v) fprintf(OUT, "%s\t%d\t%d\n", tag, p.first, p.second);
}

static void dump_sclips(const char* stage, int end, robin_hood::unordered_map<int, Sclip*>* sc) {
  std::vector<int> pos;
  for (auto& p : *sc) pos.push_back(p.first);
  std::sort(pos.begin(), pos.end());
  for (int p : pos) {
    Sclip* s = sc->at(p);
    fprintf(OUT, "%s.SC%d\t%d\t", stage, end, p);
    dump_variation_fields(s);
    fprintf(OUT, "\t%d\n", s->used ? 1 : 0);
    for (auto& ie : s->nt) {
      std::vector<std::pair<char, int> > b;
      for (auto& e : ie.second) b.push_back(std::make_pair(e.first, e.second));
      std::sort(b.begin(), b.end());
      for (auto& bc : b) fprintf(OUT, "%s.SCNT\t%d\t%d\t%d\t%c\t%d\n", stage, end, p, ie.first, bc.first, bc.second);
    }
    for (auto& ie : s->seq) {
      std::vector<char> b;
      for (auto& kv : ie.second) b.push_back(kv.first);
      std::sort(b.begin(), b.end());
      for (char c : b) {
        fprintf(OUT, "%s.SCSEQ\t%d\t%d\t%d\t%c\t", stage, end, p, ie.first, c);
        dump_variation_fields(ie.second.at(c));
        fputc('\n', OUT);
      }
    }
  }
}

static void dump_count_map(const char* tag, robin_hood::unordered_map<int, robin_hood::unordered_map<string, int> >* m) {
  std::vector<std::pair<std::pair<int, std::string>, int> > v;
  for (auto& p : *m)
    for (auto& k : p.second) v.push_back(std::make_pair(std::make_pair(p.first, k.first), k.second));
  std::sort(v.begin(), v.end());
  for (auto& e : v) fprintf(OUT, "%s\t%d\t%s\t%d\n", tag, e.first.first, e.first.second.c_str(), e.second);
}

static void dump_variant(const char* tag, int pos, const Variant* v) {
  fprintf(OUT, "%s\t%d\t%s\t%d\t%d\t%d\t%s\t%.17g\t%.17g\t%d\t%.17g\t%d\t%.17g\t%.17g\t%.17g\t%.17g\t%d\t%.17g\t%d\t%.17g\t%d\t%d\t%s\t%s\t%d\t%d\t%d\t%d\t%d\t%s\t%s\t%s\n",
          tag, pos, v->descriptionString.c_str(), v->positionCoverage, v->varsCountOnForward, v->varsCountOnReverse,
          v->strandBiasFlag.c_str(), v->frequency, v->meanPosition, v->isAtLeastAt2Positions ? 1 : 0, v->meanQuality,
          v->hasAtLeast2DiffQualities ? 1 : 0, v->meanMappingQuality, v->highQualityToLowQualityRatio,
          v->highQualityReadsFrequency, v->extraFrequency, v->shift3, v->msi, v->msint, v->numberOfMismatches, v->hicnt,
          v->hicov, v->leftseq.empty() ? "." : v->leftseq.c_str(), v->rightseq.empty() ? "." : v->rightseq.c_str(),
          v->startPosition, v->endPosition, v->refReverseCoverage, v->refForwardCoverage, v->totalPosCoverage,
          v->genotype.empty() ? "." : v->genotype.c_str(), v->varallele.empty() ? "." : v->varallele.c_str(),
          v->refallele.empty() ? "." : v->refallele.c_str());
}

static void dump_aligned(robin_hood::unordered_map<int, Vars*>& av) {
  std::vector<int> pos;
  for (auto& p : av) pos.push_back(p.first);
  std::sort(pos.begin(), pos.end());
  for (int p : pos) {
    Vars* vs = av.at(p);
    // a fresh `new Variant()` placeholder (ToVarsBuilder.cpp:1093) has an empty description
    if (vs->referenceVariant != NULL && !vs->referenceVariant->descriptionString.empty())
      dump_variant("V.REF", p, vs->referenceVariant);
    for (Variant* v : vs->variants) dump_variant("V.VAR", p, v);
  }
}

static vector<bamReader> open_readers(const string& names) {
  vector<bamReader> out;
  for (string bamname : ssplit(names, ":")) {
    samFile* in = sam_open(bamname.c_str(), "r");
    if (!in) { fprintf(stderr, "ref_dump: cannot open %s\n", bamname.c_str()); exit(1); }
    bam_hdr_t* header = sam_hdr_read(in);
    hts_idx_t* idx = sam_index_load(in, bamname.c_str());
    if (!idx) { fprintf(stderr, "ref_dump: no index for %s\n", bamname.c_str()); exit(1); }
    out.emplace_back(bamReader(in, header, idx));
  }
  return out;
}

// One sample of one region; `ref_owner` non-NULL => build the reference window with it (first sample).
static Scope<AlignedVarsData>* run_sample(const char* sample, Region region, Configuration* conf, dataPool* pool,
                                          vector<bamReader>& readers, set<string>* splice, RecordPreprocessor* pre,
                                          Reference* ref, const string& bamname, int maxReadLength, InitialData* init,
                                          bool want_c, bool want_r, bool want_v) {
  fprintf(OUT, "REGION\t%s\t%s\t%d\t%d\n", sample, region.chr.c_str(), region.start, region.end);
  Scope<InitialData> initialScope(bamname, region, ref, maxReadLength, splice, readers, init);
  CigarParser cp(pre, pool);
  Scope<VariationData> svd = cp.process(initialScope);
  if (want_c) {
    VariationData* d = svd.data;
    fprintf(OUT, "C.MAXRL\t%d\n", svd.maxReadLength);
    dump_posvarmap("C.NI", d->nonInsertionVariants);
    dump_posvarmap("C.IN", d->insertionVariants);
    dump_cov("C.COV", d->refCoverage);
    dump_sclips("C", 5, d->softClips5End);
    dump_sclips("C", 3, d->softClips3End);
    dump_count_map("C.PINS", d->positionToInsertionCount);
    dump_count_map("C.PDEL", d->positionToDeletionCount);
    dump_count_map("C.MNP", d->mnp);
  }
  VariationRealigner vr(conf, pool);
  Scope<RealignedVariationData> rvd = vr.process(svd);
  if (want_r) {
    RealignedVariationData* d = rvd.data;
    fprintf(OUT, "R.MAXRL\t%d\n", rvd.maxReadLength);
    dump_posvarmap("R.NI", d->nonInsertionVariants);
    dump_posvarmap("R.IN", d->insertionVariants);
    dump_cov("R.COV", d->refCoverage);
    dump_sclips("R", 5, d->softClips5End);
    dump_sclips("R", 3, d->softClips3End);
  }
  ToVarsBuilder tb(conf);
  Scope<AlignedVarsData>* avd = tb.process(rvd);
  if (want_v) dump_aligned(avd->data->alignedVariants);
  return avd;
}

int main(int argc, char** argv) {
  const char* outp = getenv("RV_DUMP");
  const char* stages = getenv("RV_DUMP_STAGES");  // subset of "CRV", default all
  std::string st = stages ? stages : "CRV";
  bool want_c = st.find('C') != std::string::npos, want_r = st.find('R') != std::string::npos,
       want_v = st.find('V') != std::string::npos;
  OUT = fopen(outp ? outp : "ref_dump.txt", "wb");
  if (!OUT) { fprintf(stderr, "ref_dump: cannot open output\n"); return 1; }
  Configuration* conf = cmdParse(argc, argv);
  VarDictLauncher launcher;
  launcher.start(conf);
  long total = 0;
  int nreg = 0;
  for (auto& vr : launcher.segments)
    for (auto& r : vr) { total += r.end - r.start; nreg++; }
  conf->mempool_size = (int)((total / std::max(nreg, 1)) * 1.2);
  dataPool* pool = new dataPool(conf->mempool_size);
  bool somatic = conf->bam.hasBam2();
  vector<bamReader> readers1 = open_readers(conf->bam.getBam1());
  vector<bamReader> readers2;
  if (somatic) readers2 = open_readers(conf->bam.getBam2());
  for (auto& vr : launcher.segments) {
    for (auto& region : vr) {
      set<string> splice;
      pool->reset();
      InitialData* init1 = new InitialData;
      RecordPreprocessor* pre1 = new RecordPreprocessor(region, conf, readers1);
      pre1->makeReference(conf->fasta);
      Scope<AlignedVarsData>* avd = run_sample(somatic ? "T" : "S", region, conf, pool, readers1, &splice, pre1,
                                               &(pre1->reference), conf->bam.getBam1(), 0, init1, want_c, want_r, want_v);
      if (somatic) {
        // the normal pass re-uses the tumor's Reference object and maxReadLength and an un-reset pool
        // (somaticMode.cpp:104-113)
        InitialData* init2 = new InitialData;
        RecordPreprocessor* pre2 = new RecordPreprocessor(region, conf, readers2);
        Scope<AlignedVarsData>* avd2 = run_sample("N", region, conf, pool, readers2, &splice, pre2, &(pre1->reference),
                                                  conf->bam.getBam2(), avd->maxReadLength, init2, want_c, want_r, want_v);
        delete pre2;
        delete init2;
        delete avd2->data;
        delete avd2;
      }
      delete pre1;
      delete init1;
      delete avd->data;
      delete avd;
    }
  }
  fclose(OUT);
  return 0;
}
