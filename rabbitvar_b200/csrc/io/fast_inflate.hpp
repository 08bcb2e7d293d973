// fast_inflate.hpp — raw DEFLATE (RFC 1951) decoder and CRC-32 for the BGZF blocks of a BAM file.
//
// The decode threads of the loader (file_pipeline.hpp) spend their time inflating 64 KB BGZF blocks; zlib 1.3's
// inflate() delivers ~200 MB/s per core on BAM payloads, which bounds the whole files-in -> TSV-out path.  This
// decoder is written for that one job — whole blocks, input and output both in memory, sizes known up front:
//   * a 64-bit bit buffer refilled with one unaligned 8-byte load, at most once per literal/length + distance pair;
//   * one table look-up per codeword in the common case: 11-bit primary table for literal/length codes, 8-bit for
//     distance codes, second-level tables behind the longer codes; an entry carries the symbol's base value, the
//     bits to drop for it (codeword + extra bits: ONE shift of the bit buffer on the decode loop's dependency chain,
//     the extra bits are cut out of a saved copy beside it) and its number of extra bits;
//   * matches copied eight bytes at a time (byte-wise only for distances < 8);
//   * a careful loop (byte-wise refill, every store bounds-checked) for the last few hundred bytes of a block.
// CRC-32 (IEEE 802.3, the gzip polynomial) by carry-less multiplication (PCLMULQDQ folding, Gopal et al., "Fast CRC
// Computation for Generic Polynomials Using PCLMULQDQ Instruction", Intel 2009) with a table fallback.
// Both are checked against zlib in tests/test_host_io.py on every block of the synthetic BAMs and on random streams
// of every block type.  Header-only, no dependencies beyond <immintrin.h> on x86-64.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace rvio {

class FastInflate {
 public:
  // Inflates one complete raw-deflate stream.  Returns the number of bytes written, or -1 on malformed input /
  // when the output does not fit out_cap.  Never reads beyond in + in_len nor writes beyond out + out_cap.
  long inflate(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_cap) {
    in_ = in;
    in_end_ = in + in_len;
    out_begin_ = out_ = out;
    out_end_ = out + out_cap;
    bitbuf_ = 0;
    bits_ = 0;
    phantom_ = 0;
    for (;;) {
      need(3);
      const unsigned final_block = (unsigned)(bitbuf_ & 1), type = (unsigned)((bitbuf_ >> 1) & 3);
      drop(3);
      if (type == 0) {
        if (!stored_block()) return -1;
      } else if (type == 1) {
        if (!static_ready_) {
          uint8_t lens[288 + 32];
          for (int i = 0; i < 144; ++i) lens[i] = 8;
          for (int i = 144; i < 256; ++i) lens[i] = 9;
          for (int i = 256; i < 280; ++i) lens[i] = 7;
          for (int i = 280; i < 288; ++i) lens[i] = 8;
          for (int i = 0; i < 32; ++i) lens[288 + i] = 5;
          if (!build(lens, 288, LL_BITS, static_ll_, LL_SIZE, true) || !build(lens + 288, 32, D_BITS, static_d_, D_SIZE, false))
            return -1;
          static_ready_ = true;
        }
        if (!coded_block(static_ll_, static_d_)) return -1;
      } else if (type == 2) {
        if (!read_dynamic_tables()) return -1;
        if (!coded_block(ll_, d_)) return -1;
      } else {
        return -1;
      }
      if (final_block) break;
    }
    if (phantom_ > bits_) return -1;  // the stream ran past its input
    return (long)(out_ - out_begin_);
  }

 private:
  enum { LL_BITS = 11, D_BITS = 8, LL_SIZE = 2048 + 1024, D_SIZE = 256 + 512 };
  // table entry: bits 0-5 bits to drop = codeword length still to consume + extra bits (<= 28; bits 6-7 are zero, so
  // the low byte is the shift count), 8 literal, 9 sub-table pointer, 10 end of block, 11 invalid, 12-15 extra bits
  // (or sub-table index bits for a pointer); 16-31 base value / literal / sub-table start.  The precode table of
  // read_dynamic_tables has the codeword length in its low bits and no extra bits.
  enum { F_LIT = 1u << 8, F_SUB = 1u << 9, F_EOB = 1u << 10, F_BAD = 1u << 11 };

  const uint8_t *in_, *in_end_;
  uint8_t *out_, *out_begin_, *out_end_;
  uint64_t bitbuf_;
  unsigned bits_, phantom_;
  uint32_t ll_[LL_SIZE], d_[D_SIZE], static_ll_[LL_SIZE], static_d_[D_SIZE];
  bool static_ready_ = false;

  // careful refill: at least n (<= 32) bits unless the input is exhausted (missing bits read as zero; a stream that
  // needs them fails on a later check)
  inline void need(unsigned n) {
    while (bits_ < n) {
      if (in_ < in_end_) bitbuf_ |= (uint64_t)*in_++ << bits_;
      else phantom_ += 8;  // past the end of the input: zero bits, which a well-formed stream never consumes
      bits_ += 8;
    }
  }
  inline void drop(unsigned n) { bitbuf_ >>= n; bits_ -= n; }

  bool stored_block() {
    // to the byte boundary; bytes already pulled into the bit buffer are given back
    drop(bits_ & 7);
    while (bits_ >= 8) {  // (bytes the careful refill invented past the end are not given back)
      if (phantom_ >= 8) phantom_ -= 8; else --in_;
      bits_ -= 8;
    }
    bitbuf_ = 0;
    bits_ = 0;
    if (in_end_ - in_ < 4) return false;
    const unsigned len = in_[0] | (in_[1] << 8), nlen = in_[2] | (in_[3] << 8);
    in_ += 4;
    if ((len ^ 0xffffu) != nlen) return false;
    if ((size_t)(in_end_ - in_) < len || (size_t)(out_end_ - out_) < len) return false;
    memcpy(out_, in_, len);
    in_ += len;
    out_ += len;
    return true;
  }

  static inline unsigned reverse_bits(unsigned v, int n) {  // the low n (1..16) bits of v, reversed
    v = ((v & 0x5555u) << 1) | ((v >> 1) & 0x5555u);
    v = ((v & 0x3333u) << 2) | ((v >> 2) & 0x3333u);
    v = ((v & 0x0f0fu) << 4) | ((v >> 4) & 0x0f0fu);
    v = ((v & 0x00ffu) << 8) | ((v >> 8) & 0x00ffu);
    return v >> (16 - n);
  }

  // canonical Huffman decode table from code lengths (RFC 1951 3.2.2); false for an over-subscribed code.
  // An incomplete code is accepted (zlib accepts a single distance code of one bit); unused slots are invalid.
  bool build(const uint8_t* lens, int n, int tbits, uint32_t* tab, int tab_size, bool litlen) {
    static const uint16_t len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    int count[16] = {0};
    for (int i = 0; i < n; ++i) count[lens[i]]++;
    count[0] = 0;
    unsigned next_code[16];
    unsigned code = 0;
    long left = 1;
    for (int l = 1; l <= 15; ++l) {
      left = (left << 1) - count[l];
      if (left < 0) return false;  // over-subscribed
      code = (code + (unsigned)count[l - 1]) << 1;
      next_code[l] = code;
    }
    const int primary = 1 << tbits;
    // the entry of symbol s without its codeword length
    auto entry_of = [&](int s) -> uint32_t {
      if (litlen) {
        if (s < 256) return F_LIT | ((uint32_t)s << 16);
        if (s == 256) return F_EOB;
        if (s < 286) return ((uint32_t)len_extra[s - 257] << 12) | (uint32_t)len_extra[s - 257] | ((uint32_t)len_base[s - 257] << 16);
        return F_BAD;
      }
      if (s < 30) return ((uint32_t)dist_extra[s] << 12) | (uint32_t)dist_extra[s] | ((uint32_t)dist_base[s] << 16);
      return F_BAD;
    };
    // symbols ordered by codeword length, then by symbol: the order canonical codewords are handed out in
    uint16_t first[17], put[16], sorted[288 + 32];
    first[1] = 0;
    for (int l = 1; l <= 15; ++l) { first[l + 1] = (uint16_t)(first[l] + count[l]); put[l] = first[l]; }
    for (int s = 0; s < n; ++s)
      if (lens[s]) sorted[put[lens[s]]++] = (uint16_t)s;
    // Primary table by doubling: after the codewords of l bits are in place (one store each, at their bit-reversed
    // value) the 2^l entries so far are copied behind themselves; a slot no codeword reaches stays invalid.
    tab[0] = F_BAD | 1u;
    for (int l = 1; l <= tbits; ++l) {
      const int half = 1 << (l - 1);
      memcpy(tab + half, tab, (size_t)half * sizeof(uint32_t));
      unsigned c = next_code[l];
      for (int k = first[l]; k < first[l + 1]; ++k, ++c) tab[reverse_bits(c, l)] = entry_of(sorted[k]) + (uint32_t)l;
    }
    // Second-level tables: one per distinct tbits-bit prefix of the codes longer than tbits, sized by the longest code
    // that shares the prefix — lengths ascend in `sorted`, so the last pointer written for a prefix carries it.
    int sub_next = primary;  // next free slot
    for (int l = tbits + 1; l <= 15; ++l) {
      unsigned c = next_code[l];
      for (int k = first[l]; k < first[l + 1]; ++k, ++c)
        tab[reverse_bits(c >> (l - tbits), tbits)] = F_SUB | ((uint32_t)(l - tbits) << 12) | (uint32_t)tbits;  // (start 0 = not placed yet)
    }
    for (int l = tbits + 1; l <= 15; ++l) {
      unsigned c = next_code[l];
      for (int k = first[l]; k < first[l + 1]; ++k, ++c) {
        const unsigned prefix = reverse_bits(c >> (l - tbits), tbits);
        uint32_t pe = tab[prefix];
        const int sb = (int)((pe >> 12) & 15);
        if ((pe >> 16) == 0) {
          const int sz = 1 << sb;
          if (sub_next + sz > tab_size) return false;
          pe |= (uint32_t)sub_next << 16;
          tab[prefix] = pe;
          for (int i = 0; i < sz; ++i) tab[sub_next + i] = F_BAD | 1u;
          sub_next += sz;
        }
        const int start = (int)(pe >> 16), rl = l - tbits;
        const uint32_t e = entry_of(sorted[k]) + (uint32_t)rl;  // (bits to drop: extra bits + the rest of the codeword)
        const unsigned r = reverse_bits(c & ((1u << rl) - 1u), rl);
        for (unsigned i = r; i < (1u << sb); i += 1u << rl) tab[start + i] = e;
      }
    }
    return true;
  }

  bool read_dynamic_tables() {
    need(14);
    const unsigned hlit = (unsigned)(bitbuf_ & 31) + 257, hdist = (unsigned)((bitbuf_ >> 5) & 31) + 1, hclen = (unsigned)((bitbuf_ >> 10) & 15) + 4;
    drop(14);
    if (hlit > 286 || hdist > 30) return false;
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint8_t cl[19] = {0};
    for (unsigned i = 0; i < hclen; ++i) {
      need(3);
      cl[order[i]] = (uint8_t)(bitbuf_ & 7);
      drop(3);
    }
    uint32_t ct[128 + 8];
    if (!build_precode(cl, ct)) return false;
    uint8_t lens[286 + 30 + 138];
    unsigned i = 0;
    const unsigned total = hlit + hdist;
    while (i < total) {
      need(7 + 7);
      const uint32_t e = ct[bitbuf_ & 127];
      if (e & F_BAD) return false;
      drop(e & 15);
      const unsigned sym = e >> 16;
      if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
      unsigned rep, val = 0;
      if (sym == 16) {
        if (i == 0) return false;
        val = lens[i - 1];
        rep = 3 + (unsigned)(bitbuf_ & 3);
        drop(2);
      } else if (sym == 17) {
        rep = 3 + (unsigned)(bitbuf_ & 7);
        drop(3);
      } else {
        rep = 11 + (unsigned)(bitbuf_ & 127);
        drop(7);
      }
      if (i + rep > total) return false;
      while (rep--) lens[i++] = (uint8_t)val;
    }
    if (lens[256] == 0) return false;  // no end-of-block code
    return build(lens, (int)hlit, LL_BITS, ll_, LL_SIZE, true) && build(lens + hlit, (int)hdist, D_BITS, d_, D_SIZE, false);
  }
  bool build_precode(const uint8_t* cl, uint32_t* tab) {
    int count[8] = {0};
    for (int i = 0; i < 19; ++i) count[cl[i]]++;
    count[0] = 0;
    unsigned next_code[8], code = 0;
    long left = 1;
    for (int l = 1; l <= 7; ++l) {
      left = (left << 1) - count[l];
      if (left < 0) return false;
      code = (code + (unsigned)count[l - 1]) << 1;
      next_code[l] = code;
    }
    for (int i = 0; i < 128; ++i) tab[i] = F_BAD | 1u;
    for (int s = 0; s < 19; ++s) {
      const int l = cl[s];
      if (!l) continue;
      const unsigned r = reverse_bits(next_code[l]++, l);
      for (unsigned i = r; i < 128; i += 1u << l) tab[i] = ((uint32_t)s << 16) | (uint32_t)l;
    }
    return true;
  }

  static inline uint64_t load64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
  static inline void store64(uint8_t* p, uint64_t v) { memcpy(p, &v, 8); }

  // the extra bits of a symbol whose entry is e, out of the bit buffer as it was before the symbol was dropped
  static inline unsigned extra_of(uint64_t saved, uint32_t e) {
    const unsigned tot = e & 63u, xb = (e >> 12) & 15u;
    return (unsigned)((saved & (((uint64_t)1 << tot) - 1u)) >> (tot - xb));
  }

  bool coded_block(const uint32_t* lt, const uint32_t* dt) {
    const uint64_t ll_mask = (1u << LL_BITS) - 1, d_mask = (1u << D_BITS) - 1;
    // ---- fast loop: >= 16 input bytes and >= 280 output bytes of slack, so neither needs a check inside ----
    while (in_end_ - in_ >= 16 && out_end_ - out_ >= 280) {
      // refill to >= 56 bits: a literal/length codeword (15) + extra (5) + distance codeword (15) + extra (13) = 48
      bitbuf_ |= load64(in_) << bits_;
      in_ += (63 - bits_) >> 3;
      bits_ |= 56;
      uint32_t e = lt[bitbuf_ & ll_mask];
      if (e & F_SUB) {
        drop(LL_BITS);
        e = lt[(e >> 16) + (bitbuf_ & ((1u << ((e >> 12) & 15)) - 1u))];
      }
      uint64_t saved = bitbuf_;
      drop(e & 63u);  // codeword and extra bits in one shift
      if (e & F_LIT) {
        *out_++ = (uint8_t)(e >> 16);
        // more literals from the same refill (primary-table hits only, <= 11 bits each)
        while (bits_ >= 11 && ((e = lt[bitbuf_ & ll_mask]) & F_LIT)) {
          drop(e & 63u);
          *out_++ = (uint8_t)(e >> 16);
        }
        continue;
      }
      if (e & (F_EOB | F_BAD)) return (e & F_EOB) != 0;
      {
        const unsigned len = (e >> 16) + extra_of(saved, e);
        uint32_t de = dt[bitbuf_ & d_mask];
        if (de & F_SUB) {
          drop(D_BITS);
          de = dt[(de >> 16) + (bitbuf_ & ((1u << ((de >> 12) & 15)) - 1u))];
        }
        if (de & F_BAD) return false;
        saved = bitbuf_;
        drop(de & 63u);
        const unsigned dist = (de >> 16) + extra_of(saved, de);
        if (dist > (size_t)(out_ - out_begin_)) return false;
        uint8_t* dst = out_;
        const uint8_t* src = out_ - dist;
        out_ += len;
        if (dist >= 8) {
          store64(dst, load64(src));
          store64(dst + 8, load64(src + 8));
          if (len > 16) {
            uint8_t* const end = dst + len;
            dst += 16; src += 16;
            do { store64(dst, load64(src)); dst += 8; src += 8; } while (dst < end);
          }
        } else if (dist == 1) {
          const uint64_t v = 0x0101010101010101ull * *src;
          uint8_t* const end = dst + len;
          do { store64(dst, v); dst += 8; } while (dst < end);
        } else {
          uint8_t* const end = dst + len;
          do { *dst++ = *src++; } while (dst < end);
        }
      }
    }
    // ---- careful loop for the tail of the block ----
    for (;;) {
      need(32);
      uint32_t e = lt[bitbuf_ & ll_mask];
      if (e & F_SUB) {
        drop(LL_BITS);
        e = lt[(e >> 16) + (bitbuf_ & ((1u << ((e >> 12) & 15)) - 1u))];
      }
      uint64_t saved = bitbuf_;
      drop(e & 63u);
      if (e & F_LIT) {
        if (out_ >= out_end_) return false;
        *out_++ = (uint8_t)(e >> 16);
        continue;
      }
      if (e & F_EOB) return true;
      if (e & F_BAD) return false;
      const unsigned len = (e >> 16) + extra_of(saved, e);
      need(32);
      uint32_t de = dt[bitbuf_ & d_mask];
      if (de & F_SUB) {
        drop(D_BITS);
        de = dt[(de >> 16) + (bitbuf_ & ((1u << ((de >> 12) & 15)) - 1u))];
      }
      if (de & F_BAD) return false;
      saved = bitbuf_;
      drop(de & 63u);
      const unsigned dist = (de >> 16) + extra_of(saved, de);
      if (dist > (size_t)(out_ - out_begin_) || len > (size_t)(out_end_ - out_)) return false;
      const uint8_t* src = out_ - dist;
      for (unsigned k = 0; k < len; ++k) out_[k] = src[k];
      out_ += len;
    }
  }
};

// ------------------------------------------------------------------------------------------------
// CRC-32 (reflected 0x04C11DB7), same value as zlib's crc32(0, buf, len)
// ------------------------------------------------------------------------------------------------
inline const uint32_t* crc32_table8() {
  static uint32_t t[8][256];
  static bool ready = false;
  if (!ready) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0xedb88320u & (0u - (c & 1u)));
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 255];
    ready = true;
  }
  return &t[0][0];
}
// state in / state out (the pre- and post-inversion are the caller's)
inline uint32_t crc32_update_table(uint32_t c, const uint8_t* p, size_t n) {
  const uint32_t(*t)[256] = (const uint32_t(*)[256])crc32_table8();
  while (n >= 8) {
    uint64_t v;
    memcpy(&v, p, 8);
    v ^= c;
    c = t[7][v & 255] ^ t[6][(v >> 8) & 255] ^ t[5][(v >> 16) & 255] ^ t[4][(v >> 24) & 255] ^ t[3][(v >> 32) & 255] ^
        t[2][(v >> 40) & 255] ^ t[1][(v >> 48) & 255] ^ t[0][v >> 56];
    p += 8;
    n -= 8;
  }
  while (n--) c = (c >> 8) ^ t[0][(c ^ *p++) & 255];
  return c;
}
#if defined(__x86_64__)
// n >= 64 and a multiple of 16
__attribute__((target("pclmul,sse4.1"))) inline uint32_t crc32_update_clmul(uint32_t c, const uint8_t* p, size_t n) {
  // fold constants x^(512+64), x^512, x^(128+64), x^128, x^64 mod P (bit-reflected), P and floor(x^64 / P)
  const __m128i k1k2 = _mm_set_epi64x(0x01c6e41596ll, 0x0154442bd4ll);
  const __m128i k3k4 = _mm_set_epi64x(0x00ccaa009ell, 0x01751997d0ll);
  const __m128i k5k0 = _mm_set_epi64x(0, 0x0163cd6124ll);
  const __m128i poly = _mm_set_epi64x(0x01f7011641ll, 0x01db710641ll);
  __m128i x1 = _mm_loadu_si128((const __m128i*)(p + 0)), x2 = _mm_loadu_si128((const __m128i*)(p + 16));
  __m128i x3 = _mm_loadu_si128((const __m128i*)(p + 32)), x4 = _mm_loadu_si128((const __m128i*)(p + 48));
  x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)c));
  p += 64;
  n -= 64;
  while (n >= 64) {
    __m128i y1 = _mm_clmulepi64_si128(x1, k1k2, 0x00), y2 = _mm_clmulepi64_si128(x2, k1k2, 0x00);
    __m128i y3 = _mm_clmulepi64_si128(x3, k1k2, 0x00), y4 = _mm_clmulepi64_si128(x4, k1k2, 0x00);
    x1 = _mm_clmulepi64_si128(x1, k1k2, 0x11); x2 = _mm_clmulepi64_si128(x2, k1k2, 0x11);
    x3 = _mm_clmulepi64_si128(x3, k1k2, 0x11); x4 = _mm_clmulepi64_si128(x4, k1k2, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, y1), _mm_loadu_si128((const __m128i*)(p + 0)));
    x2 = _mm_xor_si128(_mm_xor_si128(x2, y2), _mm_loadu_si128((const __m128i*)(p + 16)));
    x3 = _mm_xor_si128(_mm_xor_si128(x3, y3), _mm_loadu_si128((const __m128i*)(p + 32)));
    x4 = _mm_xor_si128(_mm_xor_si128(x4, y4), _mm_loadu_si128((const __m128i*)(p + 48)));
    p += 64;
    n -= 64;
  }
  __m128i y = _mm_clmulepi64_si128(x1, k3k4, 0x00);
  x1 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x1, k3k4, 0x11), x2), y);
  y = _mm_clmulepi64_si128(x1, k3k4, 0x00);
  x1 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x1, k3k4, 0x11), x3), y);
  y = _mm_clmulepi64_si128(x1, k3k4, 0x00);
  x1 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x1, k3k4, 0x11), x4), y);
  while (n >= 16) {
    y = _mm_clmulepi64_si128(x1, k3k4, 0x00);
    x1 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x1, k3k4, 0x11), _mm_loadu_si128((const __m128i*)p)), y);
    p += 16;
    n -= 16;
  }
  // 128 -> 64 -> 32 bits, then Barrett reduction
  const __m128i mask32 = _mm_setr_epi32(~0, 0, ~0, 0);
  x2 = _mm_clmulepi64_si128(x1, k3k4, 0x10);
  x1 = _mm_xor_si128(_mm_srli_si128(x1, 8), x2);
  x2 = _mm_srli_si128(x1, 4);
  x1 = _mm_and_si128(x1, mask32);
  x1 = _mm_xor_si128(_mm_clmulepi64_si128(x1, k5k0, 0x00), x2);
  x2 = _mm_and_si128(x1, mask32);
  x2 = _mm_clmulepi64_si128(x2, poly, 0x10);
  x2 = _mm_and_si128(x2, mask32);
  x2 = _mm_clmulepi64_si128(x2, poly, 0x00);
  x1 = _mm_xor_si128(x1, x2);
  return (uint32_t)_mm_extract_epi32(x1, 1);
}
#endif
inline uint32_t fast_crc32(const uint8_t* p, size_t n) {
  uint32_t c = 0xffffffffu;
#if defined(__x86_64__)
  static const bool have = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
  if (have && n >= 64) {
    const size_t m = n & ~(size_t)15;
    c = crc32_update_clmul(c, p, m);
    p += m;
    n -= m;
  }
#endif
  c = crc32_update_table(c, p, n);
  return c ^ 0xffffffffu;
}

}  // namespace rvio
