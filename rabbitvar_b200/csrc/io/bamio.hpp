// bamio.hpp — minimal BGZF / BAM / BAI / FASTA+FAI reader and writer over zlib.
//
// Host-side IO substrate of rabbitvar_b200.  htslib is not available in the build image, so the
// host decode path (BGZF inflate -> BAM record parse -> pinned SoA staging buffers) is written
// here from the SAM/BAM specification.  It replaces the htslib calls the reference makes in
// src/recordPreprocessor.cpp:10-32 (sam_itr_querys), :101-103 (sam_itr_next) and :56-65
// (fai_load / fai_fetch).  The same code backs the htslib-API shim under oracle/hts_shim that lets
// the unmodified reference compile as the parity oracle.
//
// Header-only, C++11, depends on zlib only.
#pragma once
#include <zlib.h>
#include "fast_inflate.hpp"
#include <fcntl.h>
#include <unistd.h>
#include <sys/stat.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <map>
#include <algorithm>
#include <stdexcept>
#include <memory>

namespace rvio {

// ------------------------------------------------------------------------------------------------
// BGZF
// ------------------------------------------------------------------------------------------------
static const int BGZF_MAX_BLOCK = 0x10000;
static const int BGZF_BLOCK_DATA = 0xff00;  // uncompressed payload per block (htslib's choice)

class BgzfReader {
 public:
  BgzfReader() : fp_(NULL), block_addr_(0), block_len_(0), block_off_(0), next_addr_(0), eof_(false) {}
  ~BgzfReader() { close(); }
  bool open(const std::string& path) {
    close();
    fp_ = fopen(path.c_str(), "rb");
    if (!fp_) return false;
    setvbuf(fp_, NULL, _IOFBF, 1 << 20);
    block_addr_ = next_addr_ = 0;
    block_len_ = block_off_ = 0;
    eof_ = false;
    return true;
  }
  void close() {
    if (fp_) fclose(fp_);
    fp_ = NULL;
  }
  bool is_open() const { return fp_ != NULL; }
  // virtual offset = (compressed block address << 16) | offset inside the inflated block
  uint64_t tell() const {
    if (block_off_ == block_len_ && block_len_ != 0) return (uint64_t)next_addr_ << 16;
    return ((uint64_t)block_addr_ << 16) | (uint64_t)block_off_;
  }
  bool seek(uint64_t voff) {
    uint64_t caddr = voff >> 16;
    uint32_t uoff = (uint32_t)(voff & 0xffff);
    if (block_len_ != 0 && caddr == block_addr_ && uoff <= block_len_) {
      block_off_ = uoff;
      return true;
    }
    if (fseeko(fp_, (off_t)caddr, SEEK_SET) != 0) return false;
    next_addr_ = caddr;
    eof_ = false;
    block_len_ = block_off_ = 0;
    if (!load_block()) return uoff == 0;
    if (uoff > block_len_) return false;
    block_off_ = uoff;
    return true;
  }
  // read exactly n bytes; returns bytes read (short only at EOF)
  size_t read(void* dst, size_t n) {
    uint8_t* out = (uint8_t*)dst;
    size_t got = 0;
    while (got < n) {
      if (block_off_ == block_len_) {
        if (!load_block()) break;
        if (block_len_ == 0) continue;  // empty block (e.g. EOF marker), try the next
      }
      size_t take = std::min((size_t)(block_len_ - block_off_), n - got);
      memcpy(out + got, buf_ + block_off_, take);
      block_off_ += (uint32_t)take;
      got += take;
    }
    return got;
  }

 private:
  bool load_block() {
    if (eof_) return false;
    uint8_t hdr[18];
    block_addr_ = next_addr_;
    size_t r = fread(hdr, 1, 18, fp_);
    if (r != 18) {
      eof_ = true;
      block_len_ = block_off_ = 0;
      return false;
    }
    if (hdr[0] != 0x1f || hdr[1] != 0x8b || hdr[2] != 8 || !(hdr[3] & 4))
      throw std::runtime_error("bgzf: bad block header");
    // walk the extra field for the BC subfield
    int xlen = hdr[10] | (hdr[11] << 8);
    uint8_t extra[256];
    int bsize = -1;
    // we already consumed 6 bytes of the extra area (hdr[12..17]) under the common layout
    memcpy(extra, hdr + 12, 6);
    if (xlen > 6) {
      if (xlen > 256 || fread(extra + 6, 1, xlen - 6, fp_) != (size_t)(xlen - 6))
        throw std::runtime_error("bgzf: truncated extra field");
    }
    for (int p = 0; p + 4 <= xlen;) {
      int slen = extra[p + 2] | (extra[p + 3] << 8);
      if (extra[p] == 'B' && extra[p + 1] == 'C' && slen == 2) bsize = (extra[p + 4] | (extra[p + 5] << 8)) + 1;
      p += 4 + slen;
    }
    if (bsize < 0) throw std::runtime_error("bgzf: no BC subfield");
    int remain = bsize - 12 - xlen;  // deflate data + crc32 + isize
    if (remain < 8 || remain > BGZF_MAX_BLOCK) throw std::runtime_error("bgzf: bad block size");
    if (fread(cbuf_, 1, remain, fp_) != (size_t)remain) throw std::runtime_error("bgzf: truncated block");
    next_addr_ = block_addr_ + bsize;
    uint32_t isize = cbuf_[remain - 4] | (cbuf_[remain - 3] << 8) | (cbuf_[remain - 2] << 16) |
                     ((uint32_t)cbuf_[remain - 1] << 24);
    if (isize > (uint32_t)BGZF_MAX_BLOCK) throw std::runtime_error("bgzf: bad isize");
    if (isize) {
      z_stream zs;
      memset(&zs, 0, sizeof(zs));
      zs.next_in = cbuf_;
      zs.avail_in = remain - 8;
      zs.next_out = buf_;
      zs.avail_out = BGZF_MAX_BLOCK;
      if (inflateInit2(&zs, -15) != Z_OK) throw std::runtime_error("bgzf: inflateInit2");
      int rc = inflate(&zs, Z_FINISH);
      inflateEnd(&zs);
      if (rc != Z_STREAM_END || zs.total_out != isize) throw std::runtime_error("bgzf: inflate failed");
      uint32_t crc_want = cbuf_[remain - 8] | (cbuf_[remain - 7] << 8) | (cbuf_[remain - 6] << 16) | ((uint32_t)cbuf_[remain - 5] << 24);
      if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), buf_, isize) != crc_want) throw std::runtime_error("bgzf: CRC mismatch");
    }
    block_len_ = isize;
    block_off_ = 0;
    return true;
  }
  FILE* fp_;
  uint64_t block_addr_;
  uint32_t block_len_, block_off_;
  uint64_t next_addr_;
  bool eof_;
  uint8_t buf_[BGZF_MAX_BLOCK];
  uint8_t cbuf_[BGZF_MAX_BLOCK];
};

class BgzfWriter {
 public:
  BgzfWriter() : fp_(NULL), fill_(0), caddr_(0), level_(1) {}
  ~BgzfWriter() { close(); }
  bool open(const std::string& path, int level = 1) {
    fp_ = fopen(path.c_str(), "wb");
    if (!fp_) return false;
    setvbuf(fp_, NULL, _IOFBF, 1 << 20);
    fill_ = 0;
    caddr_ = 0;
    level_ = level;
    return true;
  }
  uint64_t tell() const { return ((uint64_t)caddr_ << 16) | (uint64_t)fill_; }
  void write(const void* src, size_t n) {
    const uint8_t* p = (const uint8_t*)src;
    while (n) {
      size_t take = std::min(n, (size_t)(BGZF_BLOCK_DATA - fill_));
      memcpy(buf_ + fill_, p, take);
      fill_ += (uint32_t)take;
      p += take;
      n -= take;
      if (fill_ == (uint32_t)BGZF_BLOCK_DATA) flush_block();
    }
  }
  // keep a record inside one block when it fits (htslib does the same; makes voffsets simple)
  void reserve(size_t n) {
    if (n <= (size_t)BGZF_BLOCK_DATA && fill_ + n > (size_t)BGZF_BLOCK_DATA) flush_block();
  }
  void flush_block() {
    if (fill_ == 0) return;
    emit(buf_, fill_);
    fill_ = 0;
  }
  void close() {
    if (!fp_) return;
    flush_block();
    emit(buf_, 0);  // EOF marker block
    fclose(fp_);
    fp_ = NULL;
  }

 private:
  void emit(const uint8_t* data, uint32_t len) {
    uint8_t out[BGZF_MAX_BLOCK + 64];
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, level_, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK)
      throw std::runtime_error("bgzf: deflateInit2");
    zs.next_in = (Bytef*)data;
    zs.avail_in = len;
    zs.next_out = out + 18;
    zs.avail_out = BGZF_MAX_BLOCK - 18 - 8;
    int rc = deflate(&zs, Z_FINISH);
    if (rc != Z_STREAM_END) {
      deflateEnd(&zs);
      throw std::runtime_error("bgzf: deflate overflow");
    }
    uint32_t clen = (uint32_t)zs.total_out;
    deflateEnd(&zs);
    uint32_t bsize = clen + 18 + 8;
    static const uint8_t magic[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    memcpy(out, magic, 16);
    out[16] = (uint8_t)((bsize - 1) & 0xff);
    out[17] = (uint8_t)((bsize - 1) >> 8);
    uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), data, len);
    uint8_t* t = out + 18 + clen;
    t[0] = crc & 0xff; t[1] = (crc >> 8) & 0xff; t[2] = (crc >> 16) & 0xff; t[3] = (crc >> 24) & 0xff;
    t[4] = len & 0xff; t[5] = (len >> 8) & 0xff; t[6] = (len >> 16) & 0xff; t[7] = (len >> 24) & 0xff;
    if (fwrite(out, 1, bsize, fp_) != bsize) throw std::runtime_error("bgzf: write failed");
    caddr_ += bsize;
  }
  FILE* fp_;
  uint32_t fill_;
  uint64_t caddr_;
  int level_;
  uint8_t buf_[BGZF_MAX_BLOCK];
};

// ------------------------------------------------------------------------------------------------
// BAM
// ------------------------------------------------------------------------------------------------
struct BamHeader {
  std::string text;
  std::vector<std::string> names;
  std::vector<int32_t> lens;
  int tid_of(const std::string& n) const {
    for (size_t i = 0; i < names.size(); ++i)
      if (names[i] == n) return (int)i;
    return -1;
  }
};

// One alignment, raw: the 32 fixed bytes decoded + the variable-length tail kept as-is
// (qname, cigar u32[], 4-bit seq, qual, aux) exactly as the BAM specification lays them out.
struct BamRecord {
  int32_t tid, pos;  // pos is 0-based
  uint8_t l_qname, mapq;
  uint16_t bin, n_cigar, flag;
  int32_t l_seq, mtid, mpos, isize;
  std::vector<uint8_t> data;  // variable part
  const char* qname() const { return (const char*)data.data(); }
  const uint32_t* cigar() const { return (const uint32_t*)(data.data() + l_qname); }
  uint32_t* cigar() { return (uint32_t*)(data.data() + l_qname); }
  const uint8_t* seq() const { return data.data() + l_qname + 4 * (size_t)n_cigar; }
  const uint8_t* qual() const { return seq() + ((l_seq + 1) >> 1); }
  const uint8_t* aux() const { return qual() + l_seq; }
  size_t aux_len() const { return data.size() - (size_t)(aux() - data.data()); }
  // reference length consumed by the CIGAR (M, D, N, =, X)
  int32_t ref_len() const {
    int32_t l = 0;
    const uint32_t* c = cigar();
    for (int i = 0; i < n_cigar; ++i) {
      int op = c[i] & 0xf;
      if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) l += (int32_t)(c[i] >> 4);
    }
    return l;
  }
  int32_t end_pos() const {  // htslib bam_endpos: 0-based exclusive; pos+1 for unaligned
    int32_t l = (flag & 4) || n_cigar == 0 ? 0 : ref_len();
    return pos + (l ? l : 1);
  }
};

inline uint16_t reg2bin(int64_t beg, int64_t end) {
  --end;
  if (beg >> 14 == end >> 14) return (uint16_t)(((1 << 15) - 1) / 7 + (beg >> 14));
  if (beg >> 17 == end >> 17) return (uint16_t)(((1 << 12) - 1) / 7 + (beg >> 17));
  if (beg >> 20 == end >> 20) return (uint16_t)(((1 << 9) - 1) / 7 + (beg >> 20));
  if (beg >> 23 == end >> 23) return (uint16_t)(((1 << 6) - 1) / 7 + (beg >> 23));
  if (beg >> 26 == end >> 26) return (uint16_t)(((1 << 3) - 1) / 7 + (beg >> 26));
  return 0;
}

// Finds the integer value of aux tag `tag` (types c C s S i I); returns false if absent.
inline bool aux_get_int(const uint8_t* aux, size_t len, const char tag[2], int64_t* out, const uint8_t** where = NULL) {
  size_t p = 0;
  while (p + 3 <= len) {
    const uint8_t* t = aux + p;
    char ty = (char)t[2];
    size_t sz = 0;
    const uint8_t* v = t + 3;
    switch (ty) {
      case 'A': case 'c': case 'C': sz = 1; break;
      case 's': case 'S': sz = 2; break;
      case 'i': case 'I': case 'f': sz = 4; break;
      case 'd': sz = 8; break;
      case 'Z': case 'H': {
        const void* z = memchr(v, 0, len - (p + 3));
        if (!z) return false;  // unterminated string: malformed aux area
        sz = (size_t)((const uint8_t*)z - v) + 1;
        break;
      }
      case 'B': {
        if (p + 3 + 5 > len) return false;
        char sub = (char)v[0];
        uint32_t n;
        memcpy(&n, v + 1, 4);
        size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
        sz = 5 + es * n;
        break;
      }
      default: return false;
    }
    if (p + 3 + sz > len) return false;  // value runs past the record
    if (t[0] == (uint8_t)tag[0] && t[1] == (uint8_t)tag[1]) {
      if (where) *where = t + 2;
      switch (ty) {
        case 'c': *out = (int8_t)v[0]; return true;
        case 'C': *out = v[0]; return true;
        case 's': { int16_t x; memcpy(&x, v, 2); *out = x; return true; }
        case 'S': { uint16_t x; memcpy(&x, v, 2); *out = x; return true; }
        case 'i': { int32_t x; memcpy(&x, v, 4); *out = x; return true; }
        case 'I': { uint32_t x; memcpy(&x, v, 4); *out = x; return true; }
        default: *out = 0; return true;
      }
    }
    p += 3 + sz;
  }
  return false;
}

class BamReader {
 public:
  bool open(const std::string& path) {
    if (!bgzf_.open(path)) return false;
    char magic[4];
    if (bgzf_.read(magic, 4) != 4 || memcmp(magic, "BAM\1", 4) != 0) return false;
    int32_t l_text;
    if (bgzf_.read(&l_text, 4) != 4) return false;
    hdr_.text.resize(l_text);
    if (l_text && bgzf_.read(&hdr_.text[0], l_text) != (size_t)l_text) return false;
    int32_t n_ref;
    if (bgzf_.read(&n_ref, 4) != 4) return false;
    for (int i = 0; i < n_ref; ++i) {
      int32_t l_name, l_ref;
      if (bgzf_.read(&l_name, 4) != 4 || l_name < 1 || l_name > (1 << 20)) return false;
      std::string nm(l_name, '\0');
      if (bgzf_.read(&nm[0], l_name) != (size_t)l_name) return false;
      nm.resize(strlen(nm.c_str()));
      if (bgzf_.read(&l_ref, 4) != 4) return false;
      hdr_.names.push_back(nm);
      hdr_.lens.push_back(l_ref);
    }
    first_record_voff_ = bgzf_.tell();
    return true;
  }
  const BamHeader& header() const { return hdr_; }
  uint64_t tell() const { return bgzf_.tell(); }
  bool seek(uint64_t v) { return bgzf_.seek(v); }
  uint64_t first_record_voff() const { return first_record_voff_; }
  // returns false at EOF
  bool next(BamRecord& r) {
    int32_t bs;
    if (bgzf_.read(&bs, 4) != 4) return false;
    uint8_t core[32];
    if (bs < 32) throw std::runtime_error("bam: record shorter than its fixed part");
    if (bgzf_.read(core, 32) != 32) throw std::runtime_error("bam: truncated record");
    uint32_t u[8];
    memcpy(u, core, 32);
    r.tid = (int32_t)u[0];
    r.pos = (int32_t)u[1];
    r.l_qname = (uint8_t)(u[2] & 0xff);
    r.mapq = (uint8_t)((u[2] >> 8) & 0xff);
    r.bin = (uint16_t)(u[2] >> 16);
    r.n_cigar = (uint16_t)(u[3] & 0xffff);
    r.flag = (uint16_t)(u[3] >> 16);
    r.l_seq = (int32_t)u[4];
    r.mtid = (int32_t)u[5];
    r.mpos = (int32_t)u[6];
    r.isize = (int32_t)u[7];
    if (bs > (64 << 20)) throw std::runtime_error("bam: implausible record size");
    r.data.resize(bs - 32);
    if (bs > 32 && bgzf_.read(r.data.data(), bs - 32) != (size_t)(bs - 32))
      throw std::runtime_error("bam: truncated record");
    if (r.l_seq < 0 || (size_t)r.l_qname + 4 * (size_t)r.n_cigar + (size_t)((r.l_seq + 1) >> 1) + (size_t)r.l_seq > r.data.size())
      throw std::runtime_error("bam: record fields exceed its size");
    return true;
  }

 private:
  BgzfReader bgzf_;
  BamHeader hdr_;
  uint64_t first_record_voff_;
};

// BAI: we only need "where do I start scanning for reads overlapping [beg,end)" — the linear index
// gives the smallest virtual offset of any alignment overlapping each 16 kb window.  Scanning from
// there in file order and applying htslib's overlap test (pos < end && endpos > beg) returns exactly
// the records, in exactly the order, that sam_itr_querys/sam_itr_next return.
struct BaiIndex {
  struct Ref {
    std::vector<uint64_t> ioffset;
    uint64_t min_chunk_beg;  // smallest chunk start over all bins (fallback)
    bool has_data;
  };
  std::vector<Ref> refs;
  bool load(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    char magic[4];
    bool ok = fread(magic, 1, 4, f) == 4 && memcmp(magic, "BAI\1", 4) == 0;
    int32_t n_ref = 0;
    ok = ok && fread(&n_ref, 4, 1, f) == 1;
    refs.assign(ok ? n_ref : 0, Ref());
    for (int i = 0; ok && i < n_ref; ++i) {
      int32_t n_bin;
      ok = fread(&n_bin, 4, 1, f) == 1;
      refs[i].min_chunk_beg = ~0ull;
      refs[i].has_data = false;
      for (int b = 0; ok && b < n_bin; ++b) {
        uint32_t bin;
        int32_t n_chunk;
        ok = fread(&bin, 4, 1, f) == 1 && fread(&n_chunk, 4, 1, f) == 1;
        for (int c = 0; ok && c < n_chunk; ++c) {
          uint64_t be[2];
          ok = fread(be, 8, 2, f) == 2;
          if (ok && bin != 37450) {
            refs[i].min_chunk_beg = std::min(refs[i].min_chunk_beg, be[0]);
            refs[i].has_data = true;
          }
        }
      }
      int32_t n_intv = 0;
      ok = ok && fread(&n_intv, 4, 1, f) == 1;
      refs[i].ioffset.resize(ok ? n_intv : 0);
      if (ok && n_intv) ok = fread(refs[i].ioffset.data(), 8, n_intv, f) == (size_t)n_intv;
    }
    fclose(f);
    return ok;
  }
  // returns false when the reference has no alignments at or after `beg`
  bool start_offset(int tid, int64_t beg, uint64_t* voff) const {
    if (tid < 0 || tid >= (int)refs.size() || !refs[tid].has_data) return false;
    const Ref& r = refs[tid];
    size_t w = (size_t)(beg >> 14);
    if (r.ioffset.empty()) { *voff = r.min_chunk_beg; return true; }
    if (w >= r.ioffset.size()) w = r.ioffset.size() - 1;
    // htslib back-fills empty windows; we mirror that by walking left to the last non-zero entry
    uint64_t v = r.ioffset[w];
    while (v == 0 && w > 0) v = r.ioffset[--w];
    if (v == 0) v = r.min_chunk_beg;
    *voff = v;
    return true;
  }
};

// Writes a BAM plus its BAI.  Records must be appended coordinate-sorted.
class BamWriter {
 public:
  bool open(const std::string& path, const BamHeader& h, int level = 1) {
    path_ = path;
    hdr_ = h;
    if (!bgzf_.open(path, level)) return false;
    bgzf_.write("BAM\1", 4);
    int32_t l_text = (int32_t)h.text.size();
    bgzf_.write(&l_text, 4);
    bgzf_.write(h.text.data(), l_text);
    int32_t n_ref = (int32_t)h.names.size();
    bgzf_.write(&n_ref, 4);
    for (int i = 0; i < n_ref; ++i) {
      int32_t l_name = (int32_t)h.names[i].size() + 1;
      bgzf_.write(&l_name, 4);
      bgzf_.write(h.names[i].c_str(), l_name);
      bgzf_.write(&h.lens[i], 4);
    }
    bgzf_.flush_block();
    idx_.assign(n_ref, RefIdx());
    return true;
  }
  // `r.bin` is recomputed here.
  void append(BamRecord& r) {
    int32_t end = r.end_pos();
    r.bin = reg2bin(r.pos, end);
    int32_t bs = 32 + (int32_t)r.data.size();
    bgzf_.reserve(4 + (size_t)bs);
    uint64_t v0 = bgzf_.tell();
    uint32_t u[8];
    u[0] = (uint32_t)r.tid;
    u[1] = (uint32_t)r.pos;
    u[2] = (uint32_t)r.l_qname | ((uint32_t)r.mapq << 8) | ((uint32_t)r.bin << 16);
    u[3] = (uint32_t)r.n_cigar | ((uint32_t)r.flag << 16);
    u[4] = (uint32_t)r.l_seq;
    u[5] = (uint32_t)r.mtid;
    u[6] = (uint32_t)r.mpos;
    u[7] = (uint32_t)r.isize;
    bgzf_.write(&bs, 4);
    bgzf_.write(u, 32);
    bgzf_.write(r.data.data(), r.data.size());
    uint64_t v1 = bgzf_.tell();
    if (r.tid >= 0) {
      RefIdx& ri = idx_[r.tid];
      std::vector<std::pair<uint64_t, uint64_t> >& ch = ri.bins[r.bin];
      if (!ch.empty() && ch.back().second == v0) ch.back().second = v1;
      else ch.push_back(std::make_pair(v0, v1));
      size_t w0 = (size_t)(r.pos >> 14), w1 = (size_t)((end - 1) >> 14);
      if (ri.ioffset.size() <= w1) ri.ioffset.resize(w1 + 1, 0);
      for (size_t w = w0; w <= w1; ++w)
        if (ri.ioffset[w] == 0) ri.ioffset[w] = v0;
      ri.n_mapped++;
      if (ri.off_beg == 0) ri.off_beg = v0;
      ri.off_end = v1;
    }
  }
  void close() {
    bgzf_.close();
    FILE* f = fopen((path_ + ".bai").c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write bai");
    fwrite("BAI\1", 1, 4, f);
    int32_t n_ref = (int32_t)idx_.size();
    fwrite(&n_ref, 4, 1, f);
    for (int i = 0; i < n_ref; ++i) {
      RefIdx& ri = idx_[i];
      int32_t n_bin = (int32_t)ri.bins.size() + (ri.n_mapped ? 1 : 0);
      fwrite(&n_bin, 4, 1, f);
      for (std::map<uint32_t, std::vector<std::pair<uint64_t, uint64_t> > >::iterator it = ri.bins.begin();
           it != ri.bins.end(); ++it) {
        uint32_t bin = it->first;
        int32_t n_chunk = (int32_t)it->second.size();
        fwrite(&bin, 4, 1, f);
        fwrite(&n_chunk, 4, 1, f);
        for (size_t c = 0; c < it->second.size(); ++c) {
          fwrite(&it->second[c].first, 8, 1, f);
          fwrite(&it->second[c].second, 8, 1, f);
        }
      }
      if (ri.n_mapped) {  // htslib pseudo-bin with summary counts
        uint32_t bin = 37450;
        int32_t n_chunk = 2;
        uint64_t v[4] = {ri.off_beg, ri.off_end, ri.n_mapped, 0};
        fwrite(&bin, 4, 1, f);
        fwrite(&n_chunk, 4, 1, f);
        fwrite(v, 8, 4, f);
      }
      // back-fill empty linear-index windows with the next non-empty one to the right (htslib style)
      for (size_t w = ri.ioffset.size(); w-- > 1;)
        if (ri.ioffset[w - 1] == 0) ri.ioffset[w - 1] = ri.ioffset[w];
      int32_t n_intv = (int32_t)ri.ioffset.size();
      fwrite(&n_intv, 4, 1, f);
      if (n_intv) fwrite(ri.ioffset.data(), 8, n_intv, f);
    }
    uint64_t n_no_coor = 0;
    fwrite(&n_no_coor, 8, 1, f);
    fclose(f);
  }

 private:
  struct RefIdx {
    std::map<uint32_t, std::vector<std::pair<uint64_t, uint64_t> > > bins;
    std::vector<uint64_t> ioffset;
    uint64_t n_mapped, off_beg, off_end;
    RefIdx() : n_mapped(0), off_beg(0), off_end(0) {}
  };
  std::string path_;
  BamHeader hdr_;
  BgzfWriter bgzf_;
  std::vector<RefIdx> idx_;
};

// Region iterator equivalent to sam_itr_querys(idx, hdr, "chr:beg1-end1") + sam_itr_next.
class BamRegionIter {
 public:
  BamRegionIter() : rd_(NULL), tid_(-1), beg_(0), end_(0), done_(true) {}
  // beg0/end0: 0-based half-open
  void start(BamReader* rd, const BaiIndex* bai, int tid, int64_t beg0, int64_t end0) {
    rd_ = rd;
    tid_ = tid;
    beg_ = beg0;
    end_ = end0;
    uint64_t v;
    done_ = !(bai->start_offset(tid, beg0, &v) && rd->seek(v));
  }
  bool next(BamRecord& r) {
    while (!done_) {
      if (!rd_->next(r)) break;
      if (r.tid != tid_ || r.pos >= end_) break;
      if (r.end_pos() > beg_) return true;
    }
    done_ = true;
    return false;
  }

 private:
  BamReader* rd_;
  int tid_;
  int64_t beg_, end_;
  bool done_;
};

// ------------------------------------------------------------------------------------------------
// SpanScanner — the decode path of the product's loader threads: the same records, in the same order, as
// BamRegionIter, but without per-record copies.  Compressed bytes come from pread() in large chunks, every BGZF
// block is inflated (one z_stream per scanner, reset between blocks: BGZF blocks are independent deflate
// streams) onto the tail of a sliding buffer, and the records are handed to the callback IN PLACE.  The CRC32 of
// every block is checked.  A scanner belongs to one thread; threads share nothing but the page cache (the
// reference gives every OpenMP thread its own BAM handle: simpleMode.cpp:296-320).
// ------------------------------------------------------------------------------------------------
struct RawRecord {       // a BAM alignment record as it lies in the inflated stream
  const uint8_t* core;   // 32 fixed bytes (refID .. tlen), then the variable part
  int32_t block_size;    // bytes after the block_size word
  int32_t tid() const { int32_t v; memcpy(&v, core, 4); return v; }
  int32_t pos() const { int32_t v; memcpy(&v, core + 4, 4); return v; }
  uint8_t l_qname() const { return core[8]; }
  uint8_t mapq() const { return core[9]; }
  uint16_t n_cigar() const { uint16_t v; memcpy(&v, core + 12, 2); return v; }
  uint16_t flag() const { uint16_t v; memcpy(&v, core + 14, 2); return v; }
  int32_t l_seq() const { int32_t v; memcpy(&v, core + 16, 4); return v; }
  int32_t mtid() const { int32_t v; memcpy(&v, core + 20, 4); return v; }
  int32_t mpos() const { int32_t v; memcpy(&v, core + 24, 4); return v; }
  const uint8_t* var() const { return core + 32; }
  const uint8_t* cigar_bytes() const { return var() + l_qname(); }
  size_t tail_bytes() const { return 4 * (size_t)n_cigar() + (size_t)((l_seq() + 1) >> 1) + (size_t)l_seq(); }
  const uint8_t* aux() const { return cigar_bytes() + tail_bytes(); }
  size_t aux_len() const { return (size_t)block_size - 32 - l_qname() - tail_bytes(); }
  int32_t ref_len() const {
    int32_t l = 0;
    const uint8_t* c = cigar_bytes();
    for (int i = 0, n = n_cigar(); i < n; ++i) {
      uint32_t w;
      memcpy(&w, c + 4 * i, 4);
      const int op = (int)(w & 0xf);
      if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) l += (int32_t)(w >> 4);
    }
    return l;
  }
  int32_t end_pos() const {  // htslib bam_endpos
    const int32_t l = (flag() & 4) || n_cigar() == 0 ? 0 : ref_len();
    return pos() + (l ? l : 1);
  }
};

class SpanScanner {
 public:
  SpanScanner() : fd_(-1), zinit_(false), check_crc_(true) {}
  ~SpanScanner() { close(); }
  bool open(const std::string& path) {
    close();
    BamReader hdr_reader;  // the header (text + reference dictionary) through the plain reader
    if (!hdr_reader.open(path)) return false;
    hdr_ = hdr_reader.header();
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) return false;
    struct stat st;
    if (fstat(fd_, &st) != 0) return false;
    file_size_ = (uint64_t)st.st_size;
    memset(&zs_, 0, sizeof(zs_));
    if (inflateInit2(&zs_, -15) != Z_OK) return false;
    zinit_ = true;
    return true;
  }
  void close() {
    if (fd_ >= 0) ::close(fd_);
    fd_ = -1;
    if (zinit_) inflateEnd(&zs_);
    zinit_ = false;
  }
  const BamHeader& header() const { return hdr_; }
  void set_check_crc(bool on) { check_crc_ = on; }
  // Calls f(const RawRecord&) for every record of `tid` overlapping [beg0, end0) (0-based half-open), in file order.
  template <class F>
  void scan(const BaiIndex& bai, int tid, int64_t beg0, int64_t end0, F f) {
    uint64_t voff;
    if (!bai.start_offset(tid, beg0, &voff)) return;
    caddr_ = voff >> 16;
    cbuf_lo_ = cbuf_hi_ = caddr_;
    usize_ = 0;
    size_t cur = (size_t)(voff & 0xffff);
    bool first = true, eof = false;
    for (;;) {
      // complete records available at ubuf_[cur..]?
      while (usize_ >= cur + 4) {
        int32_t bs;
        memcpy(&bs, ubuf_.get() + cur, 4);
        if (bs < 32) throw std::runtime_error("bam: record shorter than its fixed part");
        if (bs > (64 << 20)) throw std::runtime_error("bam: implausible record size");
        if (usize_ < cur + 4 + (size_t)bs) break;
        RawRecord r;
        r.core = ubuf_.get() + cur + 4;
        r.block_size = bs;
        if (r.l_seq() < 0 || (size_t)r.l_qname() + r.tail_bytes() > (size_t)bs - 32)
          throw std::runtime_error("bam: record fields exceed its size");
        if (r.tid() != tid || r.pos() >= end0) return;
        if (r.end_pos() > beg0) f(r);
        cur += 4 + (size_t)bs;
      }
      if (eof) {
        if (usize_ > cur) throw std::runtime_error("bam: truncated record at end of file");
        return;
      }
      // drop what has been consumed, then append the next block
      if (!first && (cur > (size_t)(1 << 20) || cur == usize_)) {
        memmove(ubuf_.get(), ubuf_.get() + cur, usize_ - cur);
        usize_ -= cur;
        cur = 0;
      }
      if (!append_block(&eof)) eof = true;
      if (first) {
        first = false;
        if (cur > usize_) throw std::runtime_error("bam: index offset beyond its block");
      }
    }
  }

 private:
  // make [caddr_, caddr_ + n) available in cbuf_; returns a pointer or NULL at end of file
  const uint8_t* need(size_t n) {
    if (caddr_ < cbuf_lo_ || caddr_ + n > cbuf_hi_) {
      if (caddr_ + n > file_size_) return NULL;
      const size_t want = std::min<uint64_t>((uint64_t)(4 << 20), file_size_ - caddr_);
      cbuf_.resize(want);
      size_t got = 0;
      while (got < want) {
        const ssize_t r = pread(fd_, cbuf_.data() + got, want - got, (off_t)(caddr_ + got));
        if (r <= 0) break;
        got += (size_t)r;
      }
      if (got < n) return NULL;
      cbuf_lo_ = caddr_;
      cbuf_hi_ = caddr_ + got;
    }
    return cbuf_.data() + (caddr_ - cbuf_lo_);
  }
  bool append_block(bool* eof) {
    *eof = false;
    const uint8_t* h = need(18);
    if (!h) { *eof = true; return false; }
    if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) throw std::runtime_error("bgzf: bad block header");
    const int xlen = h[10] | (h[11] << 8);
    h = need(12 + (size_t)xlen);
    if (!h) throw std::runtime_error("bgzf: truncated extra field");
    int bsize = -1;
    for (int p = 0; p + 4 <= xlen;) {
      const uint8_t* x = h + 12 + p;
      const int slen = x[2] | (x[3] << 8);
      if (x[0] == 'B' && x[1] == 'C' && slen == 2) bsize = (x[4] | (x[5] << 8)) + 1;
      p += 4 + slen;
    }
    if (bsize < 0) throw std::runtime_error("bgzf: no BC subfield");
    const int remain = bsize - 12 - xlen;
    if (remain < 8) throw std::runtime_error("bgzf: bad block size");
    h = need((size_t)bsize);
    if (!h) throw std::runtime_error("bgzf: truncated block");
    const uint8_t* def = h + 12 + xlen;
    uint32_t crc_want, isize;
    memcpy(&crc_want, def + remain - 8, 4);
    memcpy(&isize, def + remain - 4, 4);
    if (isize > (uint32_t)BGZF_MAX_BLOCK) throw std::runtime_error("bgzf: bad isize");
    if (isize) {
      const size_t at = usize_;
      if (at + isize > ucap_) {  // (no zero-fill: the decoder overwrites every byte it reports)
        const size_t cap = std::max<size_t>(2 * ucap_, at + isize + (size_t)(4 << 20));
        std::unique_ptr<uint8_t[]> nb(new uint8_t[cap]);
        if (at) memcpy(nb.get(), ubuf_.get(), at);
        ubuf_.swap(nb);
        ucap_ = cap;
      }
      // the block decoder of fast_inflate.hpp; zlib only when it refuses a stream (never seen on a BGZF file)
      if (!fast_) fast_.reset(new FastInflate());
      if (fast_->inflate(def, (size_t)(remain - 8), ubuf_.get() + at, isize) != (long)isize) {
        inflateReset(&zs_);
        zs_.next_in = (Bytef*)def;
        zs_.avail_in = (uInt)(remain - 8);
        zs_.next_out = ubuf_.get() + at;
        zs_.avail_out = isize;
        const int rc = inflate(&zs_, Z_FINISH);
        if (rc != Z_STREAM_END || zs_.total_out != isize) throw std::runtime_error("bgzf: inflate failed");
      }
      if (check_crc_ && fast_crc32(ubuf_.get() + at, isize) != crc_want) throw std::runtime_error("bgzf: CRC mismatch");
      usize_ = at + isize;
    }
    caddr_ += (uint64_t)bsize;
    return true;
  }
  int fd_;
  uint64_t file_size_, caddr_, cbuf_lo_, cbuf_hi_;
  BamHeader hdr_;
  z_stream zs_;
  bool zinit_, check_crc_;
  std::unique_ptr<FastInflate> fast_;
  std::vector<uint8_t> cbuf_;
  std::unique_ptr<uint8_t[]> ubuf_;  // inflated stream, sliding
  size_t usize_ = 0, ucap_ = 0;
};

// ------------------------------------------------------------------------------------------------
// FASTA + FAI
// ------------------------------------------------------------------------------------------------
class Fasta {
 public:
  struct Entry { int64_t len, offset; int line_bases, line_width; };
  Fasta() : fp_(NULL) {}
  ~Fasta() { if (fp_) fclose(fp_); }
  bool open(const std::string& fa) {
    FILE* f = fopen((fa + ".fai").c_str(), "r");
    if (!f) return false;
    char name[1024];
    long long len, off;
    int lb, lw;
    while (fscanf(f, "%1023s %lld %lld %d %d", name, &len, &off, &lb, &lw) == 5) {
      Entry e;
      e.len = len; e.offset = off; e.line_bases = lb; e.line_width = lw;
      entries_[name] = e;
      order_.push_back(name);
    }
    fclose(f);
    fp_ = fopen(fa.c_str(), "rb");
    return fp_ != NULL;
  }
  bool has(const std::string& chr) const { return entries_.count(chr) != 0; }
  int64_t length(const std::string& chr) const {
    std::map<std::string, Entry>::const_iterator it = entries_.find(chr);
    return it == entries_.end() ? -1 : it->second.len;
  }
  // 1-based inclusive [beg1,end1], clipped to the contig; returned as stored (case preserved)
  bool fetch(const std::string& chr, int64_t beg1, int64_t end1, std::string* out) {
    std::map<std::string, Entry>::const_iterator it = entries_.find(chr);
    if (it == entries_.end()) return false;
    const Entry& e = it->second;
    if (beg1 < 1) beg1 = 1;
    if (end1 > e.len) end1 = e.len;
    out->clear();
    if (end1 < beg1) return true;
    int64_t b0 = beg1 - 1, n = end1 - beg1 + 1;
    int64_t fbeg = e.offset + (b0 / e.line_bases) * e.line_width + b0 % e.line_bases;
    int64_t e0 = b0 + n - 1;
    int64_t fend = e.offset + (e0 / e.line_bases) * e.line_width + e0 % e.line_bases + 1;
    std::vector<char> raw((size_t)(fend - fbeg));
    if (fseeko(fp_, (off_t)fbeg, SEEK_SET) != 0) return false;
    if (fread(raw.data(), 1, raw.size(), fp_) != raw.size()) return false;
    out->reserve((size_t)n);
    for (size_t i = 0; i < raw.size(); ++i)
      if (raw[i] != '\n' && raw[i] != '\r') out->push_back(raw[i]);
    return (int64_t)out->size() == n;
  }

 private:
  FILE* fp_;
  std::map<std::string, Entry> entries_;
  std::vector<std::string> order_;
};

}  // namespace rvio
