// synthgen — seeded synthetic FASTA(+FAI) / BAM(+BAI) / BED generator for the five benchmark
// configurations of BASELINE.json (shapes per SURVEY.md §8d).  Test + bench infrastructure; shares
// only the BGZF/BAM writer with the product.
//
//   synthgen --cfg 1 --out DIR [--len N] [--depth D] [--depth-n D] [--tile 10000] [--level 1]
//            [--seed S] [--snv-every N] [--indel-every N] [--softclip-frac F] [--chr NAME]
//
// Files written into DIR: ref.fa ref.fa.fai  S.bam(.bai) | T.bam N.bam  tiles.bed | panel.bed
// truth.tsv  meta.txt
//
// Generator rules that keep the data out of the reference's undefined corners (SURVEY Appendix A ⛔):
// NM always present and exact; no Q0/Q1 qualities; no N CIGAR ops / hard clips / P / B; no read
// starts or ends with I or D (>=3 matched bases kept at both ends of every read); no reads within
// 1300 bp of contig ends; read bases are ACGT only.
#include "../rabbitvar_b200/csrc/io/bamio.hpp"
#include <queue>
#include <random>
#include <functional>
#include <sys/stat.h>
#include <thread>

using namespace rvio;

struct Rng {  // splitmix64 / xoshiro256** : deterministic across platforms
  uint64_t s[4];
  static uint64_t sm(uint64_t& x) {
    uint64_t z = (x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
  }
  explicit Rng(uint64_t seed) { for (int i = 0; i < 4; ++i) s[i] = sm(seed); }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
  uint32_t below(uint32_t n) { return (uint32_t)(uni() * n); }
  int range(int lo, int hi) { return lo + (int)below((uint32_t)(hi - lo + 1)); }  // inclusive
  double normal() {
    double u1 = uni(), u2 = uni();
    if (u1 < 1e-300) u1 = 1e-300;
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
  }
};

static const char BASES[5] = "ACGT";

struct Variant {
  int pos;       // 1-based reference position of the first affected base (SNV/MNV/DEL: first
                 // changed/deleted base; INS: the base AFTER which the insertion goes)
  char type;     // 'S' snv/mnv, 'I', 'D'
  int len;       // mnv length / indel length
  std::string alt;  // substituted bases (S) or inserted bases (I)
  double vaf[2];    // per sample (0 = S or T, 1 = N)
};

struct Params {
  int cfg = 1;
  std::string out = ".";
  std::string chr = "";
  int64_t len = 0;
  double depth = 0, depth_n = 0;
  int tile = 10000;
  int level = 1;
  uint64_t seed = 0;
  int snv_every = 1000, indel_every = 5000;
  int max_indel = 30;
  double softclip_frac = 0.0;
  double mnv_frac = 0.0;
  double single_frac = 0.0;    // fraction of fragments written as two single-end records (-t cases)
  double unmapped_frac = 0.0;
  double nbase_frac = 0.0;     // fraction of reads that get 1-3 'N' bases (quality 2), NM counted like an aligner does
  double hardclip_frac = 0.0;  // fraction of reads that get an H op at one end  // fraction of pairs flagged 0x4 (placed but "unmapped": the POS-CIGAR key of -t)
  bool somatic = false, panel = false;
  int read_len = 150;
  int n_amplicons = 500, amp_len = 200, amp_space = 4000;
  double short_insert_frac = 0.0;  // overlapping mates (insert < 300)
};

static void mkdirs(const std::string& p) {
  for (size_t i = 1; i <= p.size(); ++i)
    if (i == p.size() || p[i] == '/') mkdir(p.substr(0, i).c_str(), 0755);
}

static std::string make_reference(int64_t len, Rng& rng) {
  std::string ref((size_t)len, 'A');
  for (int64_t i = 0; i < len; ++i) ref[(size_t)i] = BASES[rng.below(4)];
  // plant homopolymer / STR tracts over ~2% of positions
  int64_t planted = 0, target = len / 50;
  while (planted < target) {
    int unit = rng.range(1, 4), copies = rng.range(5, 20);
    int tl = unit * copies;
    if (len < 4000) break;
    int64_t at = 1500 + (int64_t)(rng.uni() * (double)(len - 3000 - tl));
    char u[4];
    for (int k = 0; k < unit; ++k) u[k] = BASES[rng.below(4)];
    for (int k = 0; k < tl; ++k) ref[(size_t)(at + k)] = u[k % unit];
    planted += tl;
  }
  return ref;
}

static void write_fasta(const std::string& dir, const std::string& chr, const std::string& ref) {
  FILE* f = fopen((dir + "/ref.fa").c_str(), "wb");
  if (!f) throw std::runtime_error("cannot write into " + dir);
  fprintf(f, ">%s\n", chr.c_str());
  long off = ftell(f);
  for (size_t i = 0; i < ref.size(); i += 60) {
    size_t n = std::min((size_t)60, ref.size() - i);
    fwrite(ref.data() + i, 1, n, f);
    fputc('\n', f);
  }
  fclose(f);
  f = fopen((dir + "/ref.fa.fai").c_str(), "wb");
  fprintf(f, "%s\t%zu\t%ld\t60\t61\n", chr.c_str(), ref.size(), off);
  fclose(f);
}

static char other_base(char b, Rng& rng) {
  char c;
  do c = BASES[rng.below(4)]; while (c == b);
  return c;
}

static std::vector<Variant> plant_variants(const Params& P, const std::string& ref, int64_t lo, int64_t hi, Rng& rng) {
  std::vector<Variant> vs;
  static const double VAF1[7] = {0.01, 0.02, 0.05, 0.1, 0.25, 0.5, 1.0};
  static const double VAF2[5] = {0.02, 0.05, 0.1, 0.25, 0.5};
  static const double VAF3[4] = {0.005, 0.01, 0.02, 0.05};
  std::function<void(Variant&)> set_vaf = [&](Variant& v) {
    if (P.somatic) {
      double r = rng.uni();
      if (r < 0.5) { v.vaf[0] = VAF2[rng.below(5)]; v.vaf[1] = 0; }            // somatic, tumor only
      else if (r < 0.85) { double g = rng.uni() < 0.6 ? 0.5 : 1.0; v.vaf[0] = g; v.vaf[1] = g; }  // germline
      else { v.vaf[1] = 0.5; v.vaf[0] = rng.uni() < 0.5 ? 1.0 : 0.0; }          // LOH
    } else if (P.panel) {
      v.vaf[0] = v.vaf[1] = VAF3[rng.below(4)];
    } else {
      v.vaf[0] = v.vaf[1] = VAF1[rng.below(7)];
    }
  };
  int64_t next_indel = lo + rng.range(1, P.indel_every);
  int64_t p = lo + rng.range(1, P.snv_every);
  int64_t last_end = 0;
  // SNVs / MNVs
  std::vector<Variant> snvs, indels;
  for (; p < hi - 50; p += rng.range(P.snv_every / 2, P.snv_every * 3 / 2)) {
    Variant v;
    v.pos = (int)p;
    v.type = 'S';
    v.len = (P.mnv_frac > 0 && rng.uni() < P.mnv_frac) ? rng.range(2, 4) : 1;
    for (int k = 0; k < v.len; ++k) v.alt.push_back(other_base(ref[(size_t)(p - 1 + k)], rng));
    set_vaf(v);
    snvs.push_back(v);
  }
  for (p = next_indel; p < hi - 100; p += rng.range(P.indel_every / 2, P.indel_every * 3 / 2)) {
    Variant v;
    v.pos = (int)p;
    bool ins = rng.uni() < 0.5;
    v.type = ins ? 'I' : 'D';
    double r = rng.uni();
    v.len = r < 0.5 ? rng.range(1, 3) : r < 0.85 ? rng.range(4, 12) : rng.range(13, P.max_indel);
    if (ins) for (int k = 0; k < v.len; ++k) v.alt.push_back(BASES[rng.below(4)]);
    set_vaf(v);
    indels.push_back(v);
  }
  // merge, drop overlaps (keep >= 6 bp between consecutive events)
  vs.insert(vs.end(), snvs.begin(), snvs.end());
  vs.insert(vs.end(), indels.begin(), indels.end());
  std::sort(vs.begin(), vs.end(), [](const Variant& a, const Variant& b) { return a.pos < b.pos; });
  std::vector<Variant> outv;
  for (size_t i = 0; i < vs.size(); ++i) {
    if (vs[i].pos < last_end + 6) continue;
    outv.push_back(vs[i]);
    last_end = vs[i].pos + (vs[i].type == 'I' ? 1 : vs[i].len);
  }
  return outv;
}

struct ReadOut {
  BamRecord rec;
};

// quality per base from {37:0.70, 30:0.15, 25:0.07, 20:0.05, 12:0.02, 5:0.01}; 16 random bits per base
static inline int qual_of(uint32_t r16) {
  if (r16 < 45875) return 37;
  if (r16 < 55706) return 30;
  if (r16 < 60293) return 25;
  if (r16 < 63570) return 20;
  if (r16 < 64881) return 12;
  return 5;
}
static void draw_quals(Rng& rng, uint8_t* q, int n) {
  for (int k = 0; k < n; k += 4) {
    uint64_t r = rng.next();
    for (int j = 0; j < 4 && k + j < n; ++j) q[k + j] = (uint8_t)qual_of((uint32_t)(r >> (16 * j)) & 0xffffu);
  }
}

static inline uint8_t nt16(char c) {
  switch (c) { case 'A': return 1; case 'C': return 2; case 'G': return 4; case 'T': return 8; default: return 15; }
}

// Builds one read starting at reference position `start` (1-based), `rl` read bases long, carrying the
// variants in `carry` (indices into vs, sorted).  Returns false if the read cannot be built.
static bool build_read(const Params& P, const std::string& ref, const std::vector<Variant>& vs, size_t vfirst,
                       const std::vector<char>& carried, int start, int rl, Rng& rng, double softclip_frac,
                       std::string* seq, std::vector<uint32_t>* cigar, int* nm, int* ref_start_out) {
  seq->clear();
  cigar->clear();
  *nm = 0;
  int rpos = start;  // next reference position to consume
  size_t vi = vfirst;
  while (vi < vs.size() && vs[vi].pos + (vs[vi].type == 'D' ? vs[vi].len : 0) < start) ++vi;
  std::vector<std::pair<char, int> > ops;  // run-length ops
  auto push = [&](char op, int n) {
    if (n <= 0) return;
    if (!ops.empty() && ops.back().first == op) ops.back().second += n;
    else ops.push_back(std::make_pair(op, n));
  };
  while ((int)seq->size() < rl) {
    if (rpos - 1 >= (int)ref.size()) return false;
    bool applied = false;
    if (vi < vs.size() && vs[vi].pos == rpos && carried[vi - vfirst]) {
      const Variant& v = vs[vi];
      int left = (int)seq->size(), remain = rl - left;
      if (v.type == 'S') {
        if (remain >= v.len) {
          for (int k = 0; k < v.len; ++k) seq->push_back(v.alt[k]);
          push('M', v.len);
          *nm += v.len;
          rpos += v.len;
          applied = true;
        }
      } else if (v.type == 'D') {
        // deletion removes ref bases [pos, pos+len); needs >=3 matched read bases on both sides
        if (left >= 3 && remain >= 3) {
          push('D', v.len);
          *nm += v.len;
          rpos += v.len;
          applied = true;
        }
      } else {  // insertion after base `pos`: emit the anchor base first
        if (left >= 2 && remain >= 1 + v.len + 3) {
          seq->push_back(ref[(size_t)(rpos - 1)]);
          push('M', 1);
          rpos += 1;
          for (int k = 0; k < v.len; ++k) seq->push_back(v.alt[k]);
          push('I', v.len);
          *nm += v.len;
          applied = true;
        }
      }
      ++vi;
    } else if (vi < vs.size() && vs[vi].pos < rpos) {
      ++vi;
      continue;
    }
    if (!applied) {
      // copy the matched run up to the next candidate variant in one go
      int run = rl - (int)seq->size();
      if (vi < vs.size() && vs[vi].pos > rpos) run = std::min(run, vs[vi].pos - rpos);
      else if (vi < vs.size()) run = 1;
      if (rpos - 1 + run > (int)ref.size()) return false;
      seq->append(ref, (size_t)(rpos - 1), (size_t)run);
      push('M', run);
      rpos += run;
    }
  }
  // sequencing errors (substitutions) on matched bases only
  {
    // error positions by geometric skipping (p = 0.002 per matched base); may accidentally revert a
    // planted SNV, NM is recomputed below
    int next_err = (int)(log(1.0 - rng.uni() * 0.999999) / log(1.0 - 0.002));
    int q = 0, mcount = 0;
    for (size_t o = 0; o < ops.size(); ++o) {
      if (ops[o].first == 'M') {
        while (next_err < mcount + ops[o].second) {
          int at = q + (next_err - mcount);
          (*seq)[at] = other_base((*seq)[at], rng);
          next_err += 1 + (int)(log(1.0 - rng.uni() * 0.999999) / log(1.0 - 0.002));
        }
        mcount += ops[o].second;
        q += ops[o].second;
      } else if (ops[o].first == 'I') {
        q += ops[o].second;
      }
    }
  }
  // exact NM: recompute mismatches over M ops + indel lengths
  {
    int q = 0, r = start, m = 0;
    for (size_t o = 0; o < ops.size(); ++o) {
      char op = ops[o].first;
      int n = ops[o].second;
      if (op == 'M') { for (int k = 0; k < n; ++k) m += ((*seq)[q + k] != ref[(size_t)(r - 1 + k)]); q += n; r += n; }
      else if (op == 'I') { m += n; q += n; }
      else if (op == 'D') { m += n; r += n; }
    }
    *nm = m;
  }
  *ref_start_out = start;
  // soft clips: convert a prefix/suffix of the first/last M into S.  Half keep the true (matching)
  // bases = mis-clipped by the "aligner"; half get random bases (true clip).
  if (softclip_frac > 0 && rng.uni() < softclip_frac) {
    bool at5 = rng.uni() < 0.5;
    int cl = rng.range(5, 40);
    bool matching = rng.uni() < 0.5;
    if (at5 && ops.front().first == 'M' && ops.front().second > cl + 5) {
      if (!matching) for (int k = 0; k < cl; ++k) (*seq)[k] = BASES[rng.below(4)];
      ops.front().second -= cl;
      ops.insert(ops.begin(), std::make_pair('S', cl));
      *ref_start_out = start + cl;
    } else if (!at5 && ops.back().first == 'M' && ops.back().second > cl + 5) {
      int L = (int)seq->size();
      if (!matching) for (int k = 0; k < cl; ++k) (*seq)[L - 1 - k] = BASES[rng.below(4)];
      ops.back().second -= cl;
      ops.push_back(std::make_pair('S', cl));
    }
    // NM counts only aligned bases: recompute
    int q = 0, r = *ref_start_out, m = 0;
    for (size_t o = 0; o < ops.size(); ++o) {
      char op = ops[o].first;
      int n = ops[o].second;
      if (op == 'S') q += n;
      else if (op == 'M') { for (int k = 0; k < n; ++k) m += ((*seq)[q + k] != ref[(size_t)(r - 1 + k)]); q += n; r += n; }
      else if (op == 'I') { m += n; q += n; }
      else if (op == 'D') { m += n; r += n; }
    }
    *nm = m;
  }
  for (size_t o = 0; o < ops.size(); ++o) {
    int opc = ops[o].first == 'M' ? 0 : ops[o].first == 'I' ? 1 : ops[o].first == 'D' ? 2 : 4;
    cigar->push_back(((uint32_t)ops[o].second << 4) | (uint32_t)opc);
  }
  return true;
}

// Optional extras of the edge-case data sets (drawn only when the options are set): N read bases and hard clips.
static void add_edge_cases(const Params& P, Rng& rng, const std::string& ref, int start1, std::string& seq,
                           std::vector<uint8_t>& qual, std::vector<uint32_t>& cigar, int* nm) {
  if (P.nbase_frac > 0 && rng.uni() < P.nbase_frac) {
    int k = rng.range(1, 3);
    for (int t = 0; t < k; ++t) {
      int at = (int)rng.below((uint64_t)seq.size());
      if (seq[(size_t)at] == 'N') continue;
      int q = 0, r = start1;  // which op holds read offset `at`?
      for (size_t o = 0; o < cigar.size(); ++o) {
        int op = (int)(cigar[o] & 15), n = (int)(cigar[o] >> 4);
        if (op == 0) {
          if (at < q + n) { if (seq[(size_t)at] == ref[(size_t)(r - 1 + (at - q))]) ++*nm; break; }
          q += n; r += n;
        } else if (op == 1 || op == 4) { if (at < q + n) break; q += n; }
        else if (op == 2) r += n;
      }
      seq[(size_t)at] = 'N';
      qual[(size_t)at] = 2;
    }
  }
  if (P.hardclip_frac > 0 && rng.uni() < P.hardclip_frac) {
    uint32_t h = ((uint32_t)rng.range(5, 60) << 4) | 5u;
    if (rng.uni() < 0.5) cigar.insert(cigar.begin(), h);
    else cigar.push_back(h);
  }
}

static void fill_record(BamRecord& r, const std::string& name, int tid, int pos1, int mapq, int flag,
                        const std::vector<uint32_t>& cigar, const std::string& seq, const std::vector<uint8_t>& qual,
                        int mpos1, int isize, int nm) {
  r.tid = tid;
  r.pos = pos1 - 1;
  r.l_qname = (uint8_t)(name.size() + 1);
  r.mapq = (uint8_t)mapq;
  r.n_cigar = (uint16_t)cigar.size();
  r.flag = (uint16_t)flag;
  r.l_seq = (int32_t)seq.size();
  r.mtid = tid;
  r.mpos = mpos1 - 1;
  r.isize = isize;
  r.data.clear();
  r.data.insert(r.data.end(), name.begin(), name.end());
  r.data.push_back(0);
  const uint8_t* cp = (const uint8_t*)cigar.data();
  r.data.insert(r.data.end(), cp, cp + 4 * cigar.size());
  for (size_t i = 0; i < seq.size(); i += 2) {
    uint8_t hi = nt16(seq[i]), lo = i + 1 < seq.size() ? nt16(seq[i + 1]) : 0;
    r.data.push_back((uint8_t)(hi << 4 | lo));
  }
  r.data.insert(r.data.end(), qual.begin(), qual.end());
  r.data.push_back('N'); r.data.push_back('M'); r.data.push_back('C'); r.data.push_back((uint8_t)nm);
}

struct Pending {
  BamRecord rec;
  uint64_t ord;
};
struct PendingCmp {
  bool operator()(const Pending& a, const Pending& b) const {
    if (a.rec.pos != b.rec.pos) return a.rec.pos > b.rec.pos;
    return a.ord > b.ord;
  }
};

// Generates one sample's BAM.  Fragment starts are drawn in increasing order; mates are held in a
// min-heap until the stream position passes them, so the output is coordinate-sorted.
static void make_bam(const Params& P, const std::string& path, const std::string& chr, const std::string& ref,
                     const std::vector<Variant>& vs, int sample, double depth, uint64_t seed,
                     const std::vector<std::pair<int, int> >& windows, uint64_t* n_reads_out, uint64_t* n_bases_out) {
  Rng rng(seed);
  BamHeader h;
  h.text = "@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:" + chr + "\tLN:" + std::to_string(ref.size()) + "\n";
  h.names.push_back(chr);
  h.lens.push_back((int32_t)ref.size());
  BamWriter w;
  if (!w.open(path, h, P.level)) throw std::runtime_error("cannot open " + path);
  std::priority_queue<Pending, std::vector<Pending>, PendingCmp> heap;
  uint64_t ord = 0, n_reads = 0, n_bases = 0;
  const int RL = P.read_len;
  std::string seq1, seq2;
  std::vector<uint32_t> cg1, cg2;
  std::vector<uint8_t> q1, q2;
  std::vector<char> carried;
  size_t vlo = 0;
  auto flush_to = [&](int pos0) {
    while (!heap.empty() && heap.top().rec.pos <= pos0) {
      Pending p = std::move(const_cast<Pending&>(heap.top()));
      heap.pop();
      w.append(p.rec);
    }
  };
  for (size_t wi = 0; wi < windows.size(); ++wi) {
    int wlo = windows[wi].first, whi = windows[wi].second;  // fragment start range (1-based, inclusive)
    // expected fragments so that depth ~= reads*RL/span
    double span = P.panel ? (double)P.amp_len : (double)(whi - wlo + 1);
    double nfrag = depth * span / (2.0 * RL);
    double step = (double)(whi - wlo + 1) / nfrag;
    double fpos = wlo + rng.uni() * step;
    for (; fpos <= whi; fpos += -log(1.0 - rng.uni() * 0.999999) * step) {
      int fstart = (int)fpos;
      int isz;
      if (P.short_insert_frac > 0 && rng.uni() < P.short_insert_frac) isz = rng.range(160, 299);
      else {
        isz = (int)(350 + 50 * rng.normal());
        if (isz < 160) isz = 160;
        if (isz > 800) isz = 800;
      }
      int r2start = fstart + isz - RL;
      if (r2start + RL + 100 >= (int)ref.size() - 1200) continue;
      // which variants does this fragment carry?
      while (vlo < vs.size() && vs[vlo].pos + 64 < fstart) ++vlo;
      size_t vhi = vlo;
      while (vhi < vs.size() && vs[vhi].pos <= fstart + isz + 100) ++vhi;
      carried.assign(vhi - vlo, 0);
      for (size_t k = vlo; k < vhi; ++k) carried[k - vlo] = rng.uni() < vs[k].vaf[sample] ? 1 : 0;
      int nm1, nm2, s1, s2;
      if (!build_read(P, ref, vs, vlo, carried, fstart, RL, rng, P.softclip_frac, &seq1, &cg1, &nm1, &s1)) continue;
      if (!build_read(P, ref, vs, vlo, carried, r2start, RL, rng, P.softclip_frac, &seq2, &cg2, &nm2, &s2)) continue;
      q1.resize(RL); q2.resize(RL);
      draw_quals(rng, q1.data(), RL);
      draw_quals(rng, q2.data(), RL);
      int mapq = rng.uni() < 0.9 ? 60 : rng.range(20, 59);
      bool first_fwd = rng.uni() < 0.5;  // which of read1/read2 is the forward (leftmost) mate
      int extra = 0;
      double fr = rng.uni();
      if (fr < 0.01) extra = 0x400;
      else if (fr < 0.015) extra = 0x100;
      else if (fr < 0.02) extra = 0x800;
      char nm[48];
      snprintf(nm, sizeof nm, "f%llu", (unsigned long long)ord);
      int flagL = 0x1 | 0x2 | 0x20 | (first_fwd ? 0x40 : 0x80) | extra;
      int flagR = 0x1 | 0x2 | 0x10 | (first_fwd ? 0x80 : 0x40) | extra;
      // -t cases: single-end records (no mate: RNEXT "*", PNEXT 0 => the POS-RNEXT-PNEXT duplicate key applies) and
      // paired records flagged unmapped (0x4; reach the POS-CIGAR key when -F lets them through).  The extra draws
      // happen only when the options are set, so that every other data set keeps its bytes.
      if (P.nbase_frac > 0 || P.hardclip_frac > 0) {
        add_edge_cases(P, rng, ref, s1, seq1, q1, cg1, &nm1);
        add_edge_cases(P, rng, ref, s2, seq2, q2, cg2, &nm2);
      }
      bool single = false;
      if (P.single_frac > 0 && rng.uni() < P.single_frac) single = true;
      if (!single && P.unmapped_frac > 0 && rng.uni() < P.unmapped_frac) { flagL |= 0x4; flagR |= 0x4; }
      if (single) { flagL = extra; flagR = 0x10 | extra; }
      Pending a, b;
      fill_record(a.rec, nm, 0, s1, mapq, flagL, cg1, seq1, q1, s2, isz, nm1);
      fill_record(b.rec, nm, 0, s2, mapq, flagR, cg2, seq2, q2, s1, -isz, nm2);
      if (single) {
        a.rec.mtid = b.rec.mtid = -1;
        a.rec.mpos = b.rec.mpos = -1;
        a.rec.isize = b.rec.isize = 0;
      }
      a.ord = ord * 2; b.ord = ord * 2 + 1;
      ++ord;
      flush_to(fstart - 2);  // every later read has 1-based pos >= fstart
      heap.push(std::move(a));
      heap.push(std::move(b));
      n_reads += 2;
      n_bases += 2 * RL;
    }
  }
  flush_to(1 << 30);
  w.close();
  *n_reads_out = n_reads;
  *n_bases_out = n_bases;
}

int main(int argc, char** argv) {
  Params P;
  std::map<std::string, std::string> kv;
  for (int i = 1; i + 1 < argc; i += 2) kv[argv[i]] = argv[i + 1];
  if (kv.count("--cfg")) P.cfg = atoi(kv["--cfg"].c_str());
  // presets
  switch (P.cfg) {
    case 1: P.chr = "chrS1"; P.len = 1002600; P.depth = 100; break;
    case 2: P.chr = "chrS2"; P.len = 5002600; P.depth = 200; P.depth_n = 100; P.somatic = true; break;
    case 3: P.chr = "chrS3"; P.len = 2002600; P.depth = 5000; P.panel = true; P.snv_every = 400; P.indel_every = 1500; break;
    case 4: P.chr = "chrS4"; P.len = 50002600; P.depth = 30; break;
    case 5: P.chr = "chrS5"; P.len = 5002600; P.depth = 100; P.indel_every = 500; P.max_indel = 45;
            P.softclip_frac = 0.15; P.mnv_frac = 0.3; P.short_insert_frac = 0.3; break;
    default: fprintf(stderr, "unknown --cfg\n"); return 2;
  }
  P.seed = 20261017ull + (uint64_t)P.cfg;
  if (kv.count("--out")) P.out = kv["--out"];
  if (kv.count("--len")) P.len = atoll(kv["--len"].c_str());
  if (kv.count("--depth")) P.depth = atof(kv["--depth"].c_str());
  if (kv.count("--depth-n")) P.depth_n = atof(kv["--depth-n"].c_str());
  if (kv.count("--tile")) P.tile = atoi(kv["--tile"].c_str());
  if (kv.count("--level")) P.level = atoi(kv["--level"].c_str());
  if (kv.count("--seed")) P.seed = strtoull(kv["--seed"].c_str(), NULL, 10);
  if (kv.count("--snv-every")) P.snv_every = atoi(kv["--snv-every"].c_str());
  if (kv.count("--indel-every")) P.indel_every = atoi(kv["--indel-every"].c_str());
  if (kv.count("--softclip-frac")) P.softclip_frac = atof(kv["--softclip-frac"].c_str());
  if (kv.count("--mnv-frac")) P.mnv_frac = atof(kv["--mnv-frac"].c_str());
  if (kv.count("--single-frac")) P.single_frac = atof(kv["--single-frac"].c_str());
  if (kv.count("--unmapped-frac")) P.unmapped_frac = atof(kv["--unmapped-frac"].c_str());
  if (kv.count("--nbase-frac")) P.nbase_frac = atof(kv["--nbase-frac"].c_str());
  if (kv.count("--hardclip-frac")) P.hardclip_frac = atof(kv["--hardclip-frac"].c_str());
  if (kv.count("--amplicons")) P.n_amplicons = atoi(kv["--amplicons"].c_str());
  if (kv.count("--chr")) P.chr = kv["--chr"];
  mkdirs(P.out);

  Rng rng(P.seed);
  std::string ref = make_reference(P.len, rng);
  write_fasta(P.out, P.chr, ref);
  const int64_t lo = 1301, hi = P.len - 1300;

  // regions + fragment windows
  std::vector<std::pair<int, int> > windows;
  FILE* bed = fopen((P.out + (P.panel ? "/panel.bed" : "/tiles.bed")).c_str(), "wb");
  if (P.panel) {
    int64_t at = lo + 1000;
    for (int a = 0; a < P.n_amplicons && at + P.amp_len + 1000 < hi; ++a, at += P.amp_space) {
      fprintf(bed, "%s\t%lld\t%lld\tamp%03d\n", P.chr.c_str(), (long long)at, (long long)(at + P.amp_len - 1), a);
      windows.push_back(std::make_pair((int)at - 60, (int)at + 60));
    }
  } else {
    int t = 0;
    for (int64_t s = lo; s <= hi; s += P.tile, ++t) {
      int64_t e = std::min(s + P.tile - 1, hi);
      fprintf(bed, "%s\t%lld\t%lld\ttile%05d\n", P.chr.c_str(), (long long)s, (long long)e, t);
    }
    windows.push_back(std::make_pair((int)lo, (int)(hi - 900)));
  }
  fclose(bed);

  std::vector<Variant> vs = plant_variants(P, ref, lo + 200, hi - 200, rng);
  FILE* tf = fopen((P.out + "/truth.tsv").c_str(), "wb");
  for (size_t i = 0; i < vs.size(); ++i)
    fprintf(tf, "%s\t%d\t%c\t%d\t%s\t%g\t%g\n", P.chr.c_str(), vs[i].pos, vs[i].type, vs[i].len,
            vs[i].alt.empty() ? "." : vs[i].alt.c_str(), vs[i].vaf[0], vs[i].vaf[1]);
  fclose(tf);

  uint64_t nr = 0, nb = 0, nr2 = 0, nb2 = 0;
  if (P.somatic) {
    std::thread tn([&]() { make_bam(P, P.out + "/N.bam", P.chr, ref, vs, 1, P.depth_n, P.seed * 31 + 2, windows, &nr2, &nb2); });
    make_bam(P, P.out + "/T.bam", P.chr, ref, vs, 0, P.depth, P.seed * 31 + 1, windows, &nr, &nb);
    tn.join();
  } else {
    make_bam(P, P.out + "/S.bam", P.chr, ref, vs, 0, P.depth, P.seed * 31 + 1, windows, &nr, &nb);
  }
  FILE* mf = fopen((P.out + "/meta.txt").c_str(), "wb");
  fprintf(mf, "cfg\t%d\nchr\t%s\nlen\t%lld\nregion_start\t%lld\nregion_end\t%lld\nreads\t%llu\nbases\t%llu\n"
              "reads_n\t%llu\nbases_n\t%llu\nvariants\t%zu\nseed\t%llu\n",
          P.cfg, P.chr.c_str(), (long long)P.len, (long long)lo, (long long)hi, (unsigned long long)nr,
          (unsigned long long)nb, (unsigned long long)nr2, (unsigned long long)nb2, vs.size(),
          (unsigned long long)P.seed);
  fclose(mf);
  fprintf(stderr, "synthgen cfg %d: %s len %lld, %llu(+%llu) reads, %zu variants -> %s\n", P.cfg, P.chr.c_str(),
          (long long)P.len, (unsigned long long)nr, (unsigned long long)nr2, vs.size(), P.out.c_str());
  return 0;
}
