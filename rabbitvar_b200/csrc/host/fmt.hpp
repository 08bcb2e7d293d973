// fmt.hpp — exact, allocation-free replacement for std::to_string(double) (= printf("%f"), 6 decimals, round half
// to even on the exact binary value) used by the TSV formatters: the reference prints ~40 doubles per line
// (print_output_variant_simple, simpleMode.cpp:66-142, somaticMode.cpp:151-309) and the C library call dominated
// the host stage.  The value m * 2^e is scaled by 10^6 in 128-bit integer arithmetic, so the rounding decision is
// taken on the exact value like glibc does.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>

namespace rvhost {

inline void append_f6(std::string& out, double x) {
  // anything unusual goes to the C library: non-finite values and magnitudes whose integer part needs more than
  // 64 bits after scaling
  if (!(fabs(x) < 1e12)) {
    char buf[400];
    int n = snprintf(buf, sizeof buf, "%f", x);
    out.append(buf, (size_t)n);
    return;
  }
  uint64_t bits;
  memcpy(&bits, &x, sizeof bits);
  const bool neg = (bits >> 63) != 0;
  const int ex = (int)((bits >> 52) & 0x7ff);
  uint64_t man = bits & ((1ull << 52) - 1);
  int e;  // value = man * 2^e
  if (ex == 0) e = -1074;
  else { man |= 1ull << 52; e = ex - 1075; }
  // scaled = round_half_even(man * 10^6 * 2^e)
  unsigned __int128 v = (unsigned __int128)man * 1000000u;  // < 2^73
  uint64_t scaled;
  if (e >= 0) {
    scaled = (uint64_t)(v << e);  // |x| < 1e12 keeps this below 2^64
  } else {
    const int s = -e;
    if (s >= 128) scaled = 0;
    else {
      const unsigned __int128 q = v >> s;
      const unsigned __int128 rem = v - (q << s);
      const unsigned __int128 half = (unsigned __int128)1 << (s - 1);
      scaled = (uint64_t)q;
      if (rem > half || (rem == half && (scaled & 1))) scaled++;
    }
  }
  const uint64_t ip = scaled / 1000000u;
  uint32_t fp = (uint32_t)(scaled % 1000000u);
  char buf[40];
  char* p = buf + sizeof buf;
  for (int k = 0; k < 6; ++k) { *--p = (char)('0' + fp % 10); fp /= 10; }
  *--p = '.';
  uint64_t t = ip;
  do { *--p = (char)('0' + t % 10); t /= 10; } while (t);
  if (neg) *--p = '-';  // printf prints "-0.000000" for negative values that round to zero as well
  out.append(p, (size_t)(buf + sizeof buf - p));
}

inline std::string f6(double x) {
  std::string s;
  append_f6(s, x);
  return s;
}

}  // namespace rvhost
