// hts_shim.cpp — htslib-API shim used ONLY to build the parity oracle (oracle/_ref).
//
// TEST INFRASTRUCTURE.  The reference (LeiHaoa/RabbitVar) links against htslib, which is not
// available in the build image and cannot be fetched.  It needs exactly these entry points
// (probed with `nm -u` on its objects): sam_open sam_close sam_hdr_read
// bam_hdr_destroy sam_index_load hts_idx_destroy sam_itr_querys sam_itr_next hts_itr_destroy
// bam_init1 bam_destroy1 bam_aux_get bam_aux2i bam_aux2Z bam_endpos seq_nt16_str fai_load
// fai_fetch fai_destroy kt_fisher_exact.  They are implemented here over the repo's own
// BGZF/BAM/BAI/FAI reader (rabbitvar_b200/csrc/io/bamio.hpp) from the SAM/BAM specification and
// the documented htslib semantics:
//   * sam_itr_querys("chr:b-e")  -> records with tid==chr, pos < e, bam_endpos > b-1, file order
//   * fai_fetch("chr:b-e")       -> 1-based inclusive, clipped to contig, malloc'ed
//   * kt_fisher_exact            -> htslib kfunc.c (un-vendored, unpinned: install.sh:40 clones
//                                   HEAD); restated from the published algorithm and pinned against
//                                   its exact integer evaluation (oracle/fisher_exact_rational.py).
#include "htslib/sam.h"
#include "htslib/faidx.h"
#include "htslib/kfunc.h"
#include "../../rabbitvar_b200/csrc/io/bamio.hpp"
#include <math.h>
#include <string>

struct htsFile {
  rvio::BamReader rd;
  std::string path;
};
struct hts_idx_t {
  rvio::BaiIndex bai;
};
struct hts_itr_t {
  int tid;
  int64_t beg, end;
  bool started;
  const hts_idx_t* idx;
  rvio::BamRegionIter it;
};
struct faidx_t {
  rvio::Fasta fa;
};

extern "C" {

const char seq_nt16_str[] = "=ACMGRSVTWYHKDBN";

samFile* sam_open(const char* fn, const char* mode) {
  (void)mode;
  htsFile* f = new htsFile();
  f->path = fn;
  if (!f->rd.open(fn)) {
    delete f;
    return NULL;
  }
  return f;
}
int sam_close(samFile* fp) {
  delete fp;
  return 0;
}
bam_hdr_t* sam_hdr_read(samFile* fp) {
  const rvio::BamHeader& h = fp->rd.header();
  bam_hdr_t* o = (bam_hdr_t*)calloc(1, sizeof(bam_hdr_t));
  o->n_targets = (int32_t)h.names.size();
  o->l_text = (uint32_t)h.text.size();
  o->text = strdup(h.text.c_str());
  o->target_len = (uint32_t*)calloc(o->n_targets ? o->n_targets : 1, sizeof(uint32_t));
  o->target_name = (char**)calloc(o->n_targets ? o->n_targets : 1, sizeof(char*));
  for (int i = 0; i < o->n_targets; ++i) {
    o->target_len[i] = (uint32_t)h.lens[i];
    o->target_name[i] = strdup(h.names[i].c_str());
  }
  return o;
}
void bam_hdr_destroy(bam_hdr_t* h) {
  if (!h) return;
  for (int i = 0; i < h->n_targets; ++i) free(h->target_name[i]);
  free(h->target_name);
  free(h->target_len);
  free(h->text);
  free(h);
}
hts_idx_t* sam_index_load(samFile* fp, const char* fn) {
  (void)fp;
  hts_idx_t* idx = new hts_idx_t();
  if (!idx->bai.load(std::string(fn) + ".bai")) {
    std::string alt(fn);
    size_t p = alt.rfind(".bam");
    if (p == std::string::npos || !idx->bai.load(alt.substr(0, p) + ".bai")) {
      delete idx;
      return NULL;
    }
  }
  return idx;
}
void hts_idx_destroy(hts_idx_t* idx) { delete idx; }

// "chr", "chr:pos", "chr:beg-end", commas ignored
static bool parse_region(const char* reg, std::string* chr, int64_t* beg1, int64_t* end1, bool* has_range) {
  std::string s(reg);
  size_t colon = s.rfind(':');
  *has_range = false;
  if (colon == std::string::npos) {
    *chr = s;
    *beg1 = 1;
    *end1 = (int64_t)1 << 40;
    return true;
  }
  *chr = s.substr(0, colon);
  std::string rest;
  for (size_t i = colon + 1; i < s.size(); ++i)
    if (s[i] != ',') rest.push_back(s[i]);
  size_t dash = rest.find('-');
  *has_range = true;
  if (dash == std::string::npos) {
    *beg1 = atoll(rest.c_str());
    *end1 = (int64_t)1 << 40;  // htslib: "chr:N" means N to the end
  } else {
    *beg1 = atoll(rest.substr(0, dash).c_str());
    std::string e = rest.substr(dash + 1);
    *end1 = e.empty() ? ((int64_t)1 << 40) : atoll(e.c_str());
  }
  return true;
}

hts_itr_t* sam_itr_querys(const hts_idx_t* idx, bam_hdr_t* hdr, const char* region) {
  std::string chr;
  int64_t b1, e1;
  bool hr;
  parse_region(region, &chr, &b1, &e1, &hr);
  int tid = -1;
  for (int i = 0; i < hdr->n_targets; ++i)
    if (chr == hdr->target_name[i]) tid = i;
  if (tid < 0) return NULL;
  hts_itr_t* it = new hts_itr_t();
  it->tid = tid;
  it->beg = b1 > 0 ? b1 - 1 : 0;
  it->end = e1;
  it->started = false;
  it->idx = idx;
  return it;
}
void hts_itr_destroy(hts_itr_t* iter) { delete iter; }

int sam_itr_next(samFile* fp, hts_itr_t* itr, bam1_t* r) {
  if (!itr) return -1;
  if (!itr->started) {
    itr->it.start(&fp->rd, &itr->idx->bai, itr->tid, itr->beg, itr->end);
    itr->started = true;
  }
  rvio::BamRecord rec;
  if (!itr->it.next(rec)) return -1;
  r->core.tid = rec.tid;
  r->core.pos = rec.pos;
  r->core.bin = rec.bin;
  r->core.qual = rec.mapq;
  r->core.l_qname = rec.l_qname;
  r->core.flag = rec.flag;
  r->core.unused1 = 0;
  r->core.l_extranul = 0;
  r->core.n_cigar = rec.n_cigar;
  r->core.l_qseq = rec.l_seq;
  r->core.mtid = rec.mtid;
  r->core.mpos = rec.mpos;
  r->core.isize = rec.isize;
  uint32_t need = (uint32_t)rec.data.size();
  if (r->m_data < need + 64) {
    r->m_data = need + 64;
    r->data = (uint8_t*)realloc(r->data, r->m_data);
  }
  memcpy(r->data, rec.data.data(), need);
  r->l_data = (int)need;
  return 0;
}
bam1_t* bam_init1(void) { return (bam1_t*)calloc(1, sizeof(bam1_t)); }
void bam_destroy1(bam1_t* b) {
  if (!b) return;
  free(b->data);
  free(b);
}
uint8_t* bam_aux_get(const bam1_t* b, const char tag[2]) {
  const uint8_t* aux = bam_get_aux(b);
  int l = bam_get_l_aux(b);
  int64_t v;
  const uint8_t* where = NULL;
  if (l <= 0) return NULL;
  if (!rvio::aux_get_int(aux, (size_t)l, tag, &v, &where)) return NULL;
  return (uint8_t*)where;  // points at the type byte, as htslib does
}
int64_t bam_aux2i(const uint8_t* s) {
  char ty = (char)*s++;
  switch (ty) {
    case 'c': return (int8_t)s[0];
    case 'C': return s[0];
    case 's': { int16_t x; memcpy(&x, s, 2); return x; }
    case 'S': { uint16_t x; memcpy(&x, s, 2); return x; }
    case 'i': { int32_t x; memcpy(&x, s, 4); return x; }
    case 'I': { uint32_t x; memcpy(&x, s, 4); return x; }
    default: return 0;
  }
}
char* bam_aux2Z(const uint8_t* s) {
  if (!s) return NULL;
  if (*s == 'Z' || *s == 'H') return (char*)(s + 1);
  return NULL;
}
int32_t bam_endpos(const bam1_t* b) {
  int32_t l = 0;
  if (!(b->core.flag & BAM_FUNMAP) && b->core.n_cigar > 0) {
    const uint32_t* c = bam_get_cigar(b);
    for (uint32_t i = 0; i < b->core.n_cigar; ++i)
      if (bam_cigar_type(bam_cigar_op(c[i])) & 2) l += (int32_t)bam_cigar_oplen(c[i]);
  }
  return b->core.pos + (l ? l : 1);
}

faidx_t* fai_load(const char* fn) {
  faidx_t* f = new faidx_t();
  if (!f->fa.open(fn)) {
    delete f;
    return NULL;
  }
  return f;
}
void fai_destroy(faidx_t* fai) { delete fai; }
char* fai_fetch(const faidx_t* fai, const char* reg, int* len) {
  std::string chr;
  int64_t b1, e1;
  bool hr;
  parse_region(reg, &chr, &b1, &e1, &hr);
  std::string out;
  if (!const_cast<faidx_t*>(fai)->fa.fetch(chr, b1, e1, &out)) {
    *len = -2;
    return NULL;
  }
  *len = (int)out.size();
  char* s = (char*)malloc(out.size() + 1);
  memcpy(s, out.data(), out.size());
  s[out.size()] = 0;
  return s;
}
char* faidx_fetch_seq(const faidx_t* fai, const char* c_name, int p_beg_i, int p_end_i, int* len) {
  std::string out;
  if (!const_cast<faidx_t*>(fai)->fa.fetch(c_name, (int64_t)p_beg_i + 1, (int64_t)p_end_i + 1, &out)) {
    *len = -2;
    return NULL;
  }
  *len = (int)out.size();
  char* s = (char*)malloc(out.size() + 1);
  memcpy(s, out.data(), out.size());
  s[out.size()] = 0;
  return s;
}

// ---- Fisher exact test (published algorithm of htslib kfunc.c; see header note) -----------------
static double lbinom(int n, int k) {
  if (k == 0 || n == k) return 0;
  return lgamma(n + 1) - lgamma(k + 1) - lgamma(n - k + 1);
}
static double hypergeo(int n11, int n1_, int n_1, int n) {
  return exp(lbinom(n1_, n11) + lbinom(n - n1_, n_1 - n11) - lbinom(n, n_1));
}
typedef struct { int n11, n1_, n_1, n; double p; } hgacc_t;
static double hypergeo_acc(int n11, int n1_, int n_1, int n, hgacc_t* aux) {
  if (n1_ || n_1 || n) {
    aux->n11 = n11; aux->n1_ = n1_; aux->n_1 = n_1; aux->n = n;
  } else {  // only n11 changed
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
  aux->p = hypergeo(aux->n11, aux->n1_, aux->n_1, aux->n);
  return aux->p;
}
double kt_fisher_exact(int n11, int n12, int n21, int n22, double* _left, double* _right, double* two) {
  int i, j, max, min;
  double p, q, left, right;
  hgacc_t aux;
  int n1_ = n11 + n12, n_1 = n11 + n21, n = n11 + n12 + n21 + n22;
  max = (n_1 < n1_) ? n_1 : n1_;
  min = n1_ + n_1 - n;
  if (min < 0) min = 0;
  *two = *_left = *_right = 1.;
  if (min == max) return 1.;
  q = hypergeo_acc(n11, n1_, n_1, n, &aux);
  p = hypergeo_acc(min, 0, 0, 0, &aux);
  for (left = 0., i = min + 1; p < 0.99999999 * q && i <= max; ++i) left += p, p = hypergeo_acc(i, 0, 0, 0, &aux);
  --i;
  if (p < 1.00000001 * q) left += p;
  else --i;
  p = hypergeo_acc(max, 0, 0, 0, &aux);
  for (right = 0., j = max - 1; p < 0.99999999 * q && j >= 0; --j) right += p, p = hypergeo_acc(j, 0, 0, 0, &aux);
  ++j;
  if (p < 1.00000001 * q) right += p;
  else ++j;
  *two = left + right;
  if (*two > 1.) *two = 1.;
  if (abs(i - n11) < abs(j - n11)) right = 1. - left + q;
  else left = 1.0 - right + q;
  *_left = left;
  *_right = right;
  return q;
}

}  // extern "C"
