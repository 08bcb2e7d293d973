// simple_mode_b200.cpp — REFERENCE-SIDE BINDING, compiled.  TEST INFRASTRUCTURE: shows, and lets a GPU test check, what
// a maintainer of the reference would add to drop the B200 path in.
//
// This translation unit is compiled against the reference's OWN headers (include/modes/simpleMode.h,
// include/Configuration.h, include/Region.h, used where they lie under /root/reference) and takes the place of the
// reference's src/modes/simpleMode.cpp at link time: it defines SimpleMode::process(), the function Launcher.cpp:522-524
// calls with the parsed Configuration and the RegionBuilder's segments.  Where the reference runs its OpenMP region
// loop over one_region_run (simpleMode.cpp:320-347), this one hands the same regions to rvh_run_files()
// (include/rabbitvar_b200_host.h) and writes the returned text to conf->outFileName.  Every other object of the binary
// (cmdline parsing, RegionBuilder, BED handling, sample-name logic) is the reference's, unmodified.
//
//   make -C oracle ref_gpu        ->  oracle/_ref/RabbitVar_b200 (links rabbitvar_b200/librvgpu.so)
#include "modes/simpleMode.h"
#include "../../include/rabbitvar_b200_host.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

static rv_params params_of(Configuration* conf) {
  rv_params p;
  rv_default_params(&p);
  p.goodq = conf->goodq;
  p.freq = conf->freq;
  p.lofreq = conf->lofreq;
  p.qratio = conf->qratio;
  p.mapq = conf->mapq;
  p.bias = conf->bias;
  p.vext = conf->vext;
  p.mismatch = conf->mismatch;
  p.minr = conf->minr;
  p.min_bias_reads = conf->minBiasReads;
  p.read_pos_filter = conf->readPosFilter;
  p.minmatch = conf->minmatch;
  p.trim_bases_after = conf->trimBasesAfter;
  p.indelsize = conf->indelsize;
  p.mapping_quality = conf->mappingQuality;
  p.samfilter = (int32_t)strtol(conf->samfilter.c_str(), NULL, 16);  // recordPreprocessor.cpp:121
  p.local_realign = conf->performLocalRealignment ? 1 : 0;
  p.move3 = conf->moveIndelsTo3 ? 1 : 0;
  p.uniq_u = conf->uniqueModeAlignmentEnabled ? 1 : 0;
  p.uniq_un = conf->uniqueModeSecondInPairEnabled ? 1 : 0;
  p.dedup = conf->removeDuplicatedReads ? 1 : 0;
  p.pileup = conf->doPileup ? 1 : 0;
  p.fisher = conf->fisher ? 1 : 0;
  return p;
}

void SimpleMode::process(Configuration* conf, vector<vector<Region> >& segments) {
  std::vector<std::string> chr_s, gene_s;
  std::vector<int32_t> start, end;
  for (vector<Region>& seg : segments)
    for (Region& r : seg) {
      chr_s.push_back(r.chr);
      gene_s.push_back(r.gene);
      start.push_back(r.start);
      end.push_back(r.end);
    }
  std::vector<const char*> chr, gene;
  for (size_t i = 0; i < chr_s.size(); ++i) { chr.push_back(chr_s[i].c_str()); gene.push_back(gene_s[i].c_str()); }
  const rv_params p = params_of(conf);
  char* tsv = NULL;
  int64_t len = 0;
  const int rc = rvh_run_files(&p, conf->fasta.c_str(), conf->bam.getBam1().c_str(), NULL, conf->sample.c_str(),
                               (int32_t)chr.size(), chr.data(), start.data(), end.data(), gene.data(),
                               conf->threads > 0 ? conf->threads : 1, 1, 0, &tsv, &len, NULL);
  if (rc != 0) fprintf(stderr, "rvh_run_files: rc %d: %s\n", rc, rvh_last_error());
  if (rc != 0 && rc != 2) exit(3);
  FILE* f = fopen(conf->outFileName.c_str(), "wb");
  if (!f) { cerr << "open file: " << conf->outFileName << " error!" << endl; exit(1); }
  cout << "[info] output file name: " << conf->outFileName << endl;
  fwrite(tsv, 1, (size_t)len, f);
  fclose(f);
  rvh_free(tsv);
}

// the reference's per-region members: not used by this binding (the region loop lives behind rvh_run_files)
void SimpleMode::InitItemRepository(const int) {}
Scope<AlignedVarsData>* SimpleMode::one_region_run(Region, Configuration*, dataPool*, vector<bamReader>, set<string>*) { return NULL; }
void SimpleMode::print_output_variant_simple(const Variant*, Region&, std::string, int, std::string, bool) {}
void SimpleMode::output(Scope<AlignedVarsData>*, Configuration*) {}
