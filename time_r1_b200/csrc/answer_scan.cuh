// Answer scanner: the text side of the verifiable rewards, as byte-level code that runs on the device
// (SURVEY §8f rank 3).  It reproduces, on the UTF-8 bytes of a decoded completion, exactly what the reference does with
// Python `re` / `float()` on the decoded str:
//
//   parse_timestamp_output   main.py:122-142   re.findall(r"<answer>(.*?)</answer>", s, DOTALL)[-1]
//                                              -> re.findall(r"(\d+\.?\d*) (to|and) (\d+\.?\d*)", block, IGNORECASE)[-1]
//                                              -> float(group 1), float(group 3)
//   format_reward            main.py:234-239   re.fullmatch(r"<think>.*?</think>\s*<answer>.*?</answer>", s.strip(), DOTALL)
//   extract_answer           evaluate.py:125-151 (the filtering epochs): last span match anywhere, case-sensitive
//   extract_answer_force     src/vllm_inference/calc_difficulty.py:40-46: the first two r"\d+(?:\.\d+)?" numbers
//
// Python semantics kept: `\d` is every Unicode Nd code point (float() maps it to its decimal value), `\s` / str.strip()
// is the 29-code-point Py_UNICODE_ISSPACE set, IGNORECASE on "to|and" only ever matches ASCII letters (checked by brute
// force over all code points, tests/test_answer_scan_cpu.py), bytes that are not valid UTF-8 decode to U+FFFD (the
// tokenizer decodes with errors="replace"), which is neither digit nor space, and float() is correctly rounded
// (round-half-even) for any number of digits, including subnormals, 0 and inf.
//
// Everything here is integer / byte work.  The functions are __host__ __device__ so that tests/native/answer_scan_host.cpp
// can compile this very header with g++ and check it against CPython on the CPU; the product only ever runs it inside
// the kernels of answer_scan.cu.  `W` is the execution policy: on the device a warp (32 lanes search 32 byte offsets at a
// time and vote), in the host harness a single lane.
#pragma once
#include <stdint.h>

#include "unicode_nd_table.h"

#if defined(__CUDACC__)
#define AS_HD __host__ __device__ __forceinline__
#define AS_HD_NOINLINE __host__ __device__ __noinline__
#else
#define AS_HD inline
#define AS_HD_NOINLINE inline
#endif

namespace ascan {

#if defined(__CUDACC__)
__device__ __constant__ uint32_t d_nd_start[AS_ND_RUNS] = {AS_ND_TABLE};
#endif
static const uint32_t h_nd_start[AS_ND_RUNS] = {AS_ND_TABLE};

AS_HD const uint32_t* nd_table() {
#if defined(__CUDA_ARCH__)
  return d_nd_start;
#else
  return h_nd_start;
#endif
}

struct Char {
  uint32_t cp;  // code point (0xFFFD for a byte that does not start a valid UTF-8 sequence)
  int len;      // bytes consumed (1 for an invalid byte: the decoder resynchronises on the next byte)
};

AS_HD bool is_cont(uint8_t b) { return (b & 0xC0) == 0x80; }

// Strict UTF-8 (no overlongs, no surrogates, <= U+10FFFF) — the validity rules of CPython's decoder.  An invalid or
// truncated sequence yields U+FFFD for its first byte only; with errors="replace" CPython may fold several bytes into one
// U+FFFD, but the bytes it folds are continuation bytes, which can never start a character the scanner cares about.
AS_HD Char decode(const uint8_t* t, int p, int end) {
  const uint8_t b0 = t[p];
  if (b0 < 0x80) return Char{b0, 1};
  const Char bad{0xFFFDu, 1};
  if (b0 >= 0xC2 && b0 <= 0xDF) {
    if (p + 1 < end && is_cont(t[p + 1])) return Char{((uint32_t)(b0 & 0x1F) << 6) | (t[p + 1] & 0x3Fu), 2};
    return bad;
  }
  if (b0 >= 0xE0 && b0 <= 0xEF) {
    if (p + 2 >= end) return bad;
    const uint8_t b1 = t[p + 1], b2 = t[p + 2];
    const uint8_t lo = b0 == 0xE0 ? 0xA0 : 0x80, hi = b0 == 0xED ? 0x9F : 0xBF;
    if (b1 < lo || b1 > hi || !is_cont(b2)) return bad;
    return Char{((uint32_t)(b0 & 0x0F) << 12) | ((uint32_t)(b1 & 0x3F) << 6) | (b2 & 0x3Fu), 3};
  }
  if (b0 >= 0xF0 && b0 <= 0xF4) {
    if (p + 3 >= end) return bad;
    const uint8_t b1 = t[p + 1], b2 = t[p + 2], b3 = t[p + 3];
    const uint8_t lo = b0 == 0xF0 ? 0x90 : 0x80, hi = b0 == 0xF4 ? 0x8F : 0xBF;
    if (b1 < lo || b1 > hi || !is_cont(b2) || !is_cont(b3)) return bad;
    return Char{((uint32_t)(b0 & 0x07) << 18) | ((uint32_t)(b1 & 0x3F) << 12) | ((uint32_t)(b2 & 0x3F) << 6) | (b3 & 0x3Fu), 4};
  }
  return bad;
}

// Decimal value 0..9 of a code point Python's `\d` accepts, or -1.
AS_HD int digit_value(uint32_t cp) {
  if (cp < 0x80) return (cp >= '0' && cp <= '9') ? (int)(cp - '0') : -1;
  const uint32_t* tab = nd_table();
  int lo = 0, hi = AS_ND_RUNS - 1;  // largest start <= cp
  if (cp < tab[1]) return -1;       // tab[0] is ASCII '0'
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab[mid] <= cp) lo = mid; else hi = mid - 1;
  }
  const uint32_t d = cp - tab[lo];
  return d < 10 ? (int)d : -1;
}

// Py_UNICODE_ISSPACE: what `\s` (str pattern) and str.strip() treat as whitespace.
AS_HD bool is_space(uint32_t cp) {
  if (cp < 0x80) return (cp >= 0x09 && cp <= 0x0D) || (cp >= 0x1C && cp <= 0x20);
  return cp == 0x85 || cp == 0xA0 || cp == 0x1680 || (cp >= 0x2000 && cp <= 0x200A) || cp == 0x2028 || cp == 0x2029 ||
         cp == 0x202F || cp == 0x205F || cp == 0x3000;
}

// Byte length of the whitespace character that ends exactly at `b` inside [a, b), or 0.  A lead byte is never consumed
// as part of another sequence, so decoding from it is context-free.
AS_HD int trailing_space_len(const uint8_t* t, int a, int b) {
  const uint8_t c = t[b - 1];
  if (c < 0x80) return is_space(c) ? 1 : 0;
  for (int l = 2; l <= 3; ++l) {
    if (b - l < a) break;
    const Char ch = decode(t, b - l, b);
    if (ch.len == l && is_space(ch.cp)) return l;
  }
  return 0;
}

AS_HD bool match_at(const uint8_t* t, int p, int end, const char* pat, int plen) {
  if (p + plen > end) return false;
  for (int k = 0; k < plen; ++k)
    if (t[p + k] != (uint8_t)pat[k]) return false;
  return true;
}

// First i in [from, end - plen] with t[i, i + plen) == pat, or -1.  Cooperative: lane l tests offset base + l.  The
// patterns are ASCII, and ASCII bytes never occur inside a multi-byte UTF-8 sequence, so a byte search is a str search.
template <class W>
AS_HD int find_pat(const uint8_t* t, int from, int end, const char* pat, int plen, const W& w) {
  for (int base = from; base + plen <= end; base += W::kLanes) {
    const int p = base + w.lane();
    const uint32_t m = w.ballot(match_at(t, p, end, pat, plen));
    if (m) return base + W::first_set(m);
  }
  return -1;
}

// ---------------------------------------------------------------------------------------------------------------------
// float(): correctly rounded decimal -> binary64 for a digit string D and a power of ten.
// ---------------------------------------------------------------------------------------------------------------------

constexpr int kBigLimbs = 136;  // 4352 bits: 10^1110 (3688 bits) shifted by up to 56, see to_double_slow

struct Big {
  uint32_t v[kBigLimbs];
  int n;  // used limbs (v[n-1] != 0, or n == 0 for zero)
};

AS_HD void big_zero(Big& a) { a.n = 0; }

AS_HD void big_mul_add(Big& a, uint32_t mul, uint32_t add) {  // a = a * mul + add
  uint64_t carry = add;
  for (int i = 0; i < a.n; ++i) {
    const uint64_t x = (uint64_t)a.v[i] * mul + carry;
    a.v[i] = (uint32_t)x;
    carry = x >> 32;
  }
  if (carry && a.n < kBigLimbs) a.v[a.n++] = (uint32_t)carry;
}

AS_HD int big_bitlen(const Big& a) {
  if (a.n == 0) return 0;
  uint32_t top = a.v[a.n - 1];
  int b = 0;
  while (top) { ++b; top >>= 1; }
  return (a.n - 1) * 32 + b;
}

AS_HD void big_shl(Big& a, int s) {  // a <<= s
  if (a.n == 0 || s == 0) return;
  const int ws = s >> 5, bs = s & 31;
  int nn = a.n + ws + 1;
  if (nn > kBigLimbs) nn = kBigLimbs;
  for (int i = nn - 1; i >= 0; --i) {
    const int j = i - ws;
    uint32_t x = 0;
    if (j >= 0 && j < a.n) x = a.v[j] << bs;
    if (bs && j - 1 >= 0 && j - 1 < a.n) x |= a.v[j - 1] >> (32 - bs);
    a.v[i] = x;
  }
  a.n = nn;
  while (a.n > 0 && a.v[a.n - 1] == 0) --a.n;
}

AS_HD void big_shr1(Big& a) {
  for (int i = 0; i < a.n; ++i) {
    uint32_t x = a.v[i] >> 1;
    if (i + 1 < a.n) x |= a.v[i + 1] << 31;
    a.v[i] = x;
  }
  while (a.n > 0 && a.v[a.n - 1] == 0) --a.n;
}

AS_HD bool big_ge(const Big& a, const Big& b) {  // a >= b
  if (a.n != b.n) return a.n > b.n;
  for (int i = a.n - 1; i >= 0; --i)
    if (a.v[i] != b.v[i]) return a.v[i] > b.v[i];
  return true;
}

AS_HD void big_sub(Big& a, const Big& b) {  // a -= b, a >= b
  uint64_t borrow = 0;
  for (int i = 0; i < a.n; ++i) {
    const uint64_t x = (uint64_t)a.v[i] - (i < b.n ? b.v[i] : 0u) - borrow;
    a.v[i] = (uint32_t)x;
    borrow = (x >> 63) & 1;
  }
  while (a.n > 0 && a.v[a.n - 1] == 0) --a.n;
}

AS_HD double bits_to_double(uint64_t bits) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)bits);
#else
  union { uint64_t u; double d; } x;
  x.u = bits;
  return x.d;
#endif
}

// q * 2^(-s) (q in [2^54, 2^56), `sticky` = a non-zero remainder below q's last bit) -> nearest double, ties to even.
AS_HD double round_pack(uint64_t q, int s, bool sticky) {
  int L = 0;
  for (uint64_t x = q; x; x >>= 1) ++L;
  int drop = L - 53;                            // normal: keep 53 bits
  if (s - 1074 > drop) drop = s - 1074;         // subnormal: keep down to the 2^-1074 quantum
  if (drop > 62) return 0.0;                    // below half the smallest subnormal
  uint64_t kept = q >> drop;
  const uint64_t rem = q & ((1ull << drop) - 1), half = 1ull << (drop - 1);
  if (rem > half || (rem == half && (sticky || (kept & 1)))) ++kept;
  if (kept == 0) return 0.0;
  const int e = drop - s;                       // value = kept * 2^e
  int kb = 0;
  for (uint64_t x = kept; x; x >>= 1) ++kb;
  const int biased = e + kb - 1 + 1023;
  if (biased >= 2047) return bits_to_double(0x7FF0000000000000ull);  // float() returns inf
  if (biased >= 1) {
    const uint64_t mant = kb <= 53 ? kept << (53 - kb) : kept >> (kb - 53);
    return bits_to_double(((uint64_t)biased << 52) | (mant & 0xFFFFFFFFFFFFFull));
  }
  return bits_to_double(kept);                  // subnormal (e == -1074); kept == 2^52 is the smallest normal
}

// Digits of one regex group, re-read from the text: calls f(value) for every digit, f(-1) for the '.'.
template <class F>
AS_HD void for_each_digit(const uint8_t* t, int p, int q, F&& f) {
  while (p < q) {
    const Char ch = decode(t, p, q);
    f(ch.cp == '.' ? -1 : digit_value(ch.cp));
    p += ch.len;
  }
}

// General path: value = N * 10^e10 with N the first <= 768 significant digits (plus a sticky digit when a later one is
// non-zero: no halfway point between two doubles has more than 767 significant digits), as an exact big-integer ratio
// followed by one rounding.  Bounds: N < 10^769 (2555 bits), |e10| <= 1110.
AS_HD_NOINLINE double to_double_slow(const uint8_t* t, int p, int q) {
  Big num, den;
  big_zero(num);
  int stored = 0, e10 = 0;
  bool frac = false, tail = false;
  for_each_digit(t, p, q, [&](int v) {
    if (v < 0) { frac = true; return; }
    if (stored == 0 && v == 0) { if (frac) --e10; return; }          // leading zeros
    if (stored < 768) { big_mul_add(num, 10, (uint32_t)v); ++stored; if (frac) --e10; }
    else { if (!frac) ++e10; if (v) tail = true; }
  });
  if (stored == 0) return 0.0;
  if (tail) { big_mul_add(num, 10, 1); ++stored; --e10; }
  if (stored + e10 > 310) return bits_to_double(0x7FF0000000000000ull);  // >= 10^310 -> inf
  if (stored + e10 < -330) return 0.0;                                  // < 10^-330 -> 0
  big_zero(den);
  den.v[0] = 1; den.n = 1;
  Big& scaled = e10 >= 0 ? num : den;
  for (int k = e10 >= 0 ? e10 : -e10; k > 0;) {
    const int step = k >= 9 ? 9 : k;
    uint32_t m = 1;
    for (int i = 0; i < step; ++i) m *= 10;
    big_mul_add(scaled, m, 0);
    k -= step;
  }
  // q55 = floor(num * 2^s / den) with bitlen(num * 2^s) - bitlen(den) == 55, i.e. 2^54 <= q55 < 2^56
  const int s = 55 - (big_bitlen(num) - big_bitlen(den));
  if (s > 0) big_shl(num, s);
  big_shl(den, 55 + (s < 0 ? -s : 0));          // den * 2^55 (aligned with num), halved once per quotient bit
  uint64_t quo = 0;
  for (int i = 55; i >= 0; --i) {
    if (big_ge(num, den)) { big_sub(num, den); quo |= 1ull << i; }
    if (i) big_shr1(den);                       // exact: the 55 halvings only undo the shift above
  }
  return round_pack(quo, s, num.n != 0);        // what is left of num is the remainder: the sticky bit
}

// float(text[p:q)) for text[p:q) matching \d+\.?\d* (or \d+(\.\d+)?).
AS_HD double to_double(const uint8_t* t, int p, int q) {
  uint64_t m = 0;
  int stored = 0, e10 = 0;
  bool frac = false, exact = true;
  for_each_digit(t, p, q, [&](int v) {
    if (v < 0) { frac = true; return; }
    if (m == 0 && v == 0) { if (frac) --e10; return; }
    if (stored < 19) { m = m * 10 + (uint64_t)v; ++stored; if (frac) --e10; }
    else { if (!frac) ++e10; if (v) exact = false; }
  });
  if (m == 0) return 0.0;
  if (exact && m < (1ull << 53) && e10 >= -22 && e10 <= 22) {
    // both operands exact in binary64 -> one correctly rounded IEEE operation (Clinger's fast path)
    double p10 = 1.0;
    for (int i = 0, k = e10 < 0 ? -e10 : e10; i < k; ++i) p10 *= 10.0;   // 10^k exact for k <= 22
#if defined(__CUDA_ARCH__)
    return e10 < 0 ? __ddiv_rn((double)m, p10) : __dmul_rn((double)m, p10);
#else
    return e10 < 0 ? (double)m / p10 : (double)m * p10;
#endif
  }
  return to_double_slow(t, p, q);
}

// ---------------------------------------------------------------------------------------------------------------------
// regex restatements
// ---------------------------------------------------------------------------------------------------------------------

AS_HD int skip_digits(const uint8_t* t, int p, int end) {
  while (p < end) {
    const Char ch = decode(t, p, end);
    if (digit_value(ch.cp) < 0) break;
    p += ch.len;
  }
  return p;
}

AS_HD bool ci_eq(uint8_t b, char lower) { return (b | 0x20) == (uint8_t)lower; }  // lower in {t, o, a, n, d}

struct Span {
  int s0, s1, e0, e1;  // byte ranges of group 1 and group 3; s0 < 0: no match
};

// re.findall(r"(\d+\.?\d*) (to|and) (\d+\.?\d*)", text[bs:be), flags)[-1].  The pattern has no ambiguity a backtracking
// engine could exploit (after the greedy number the next byte must be ' ', and giving back a digit or the '.' leaves a
// digit or '.' there), so every start position has at most one way to match: scan start positions left to right, resume
// after each match.  A failed attempt fails for every start inside its first digit run as well (they reach the same
// group-1 end), so the scan jumps over the run and stays linear.
AS_HD Span last_span(const uint8_t* t, int bs, int be, bool ignore_case) {
  Span last{-1, -1, -1, -1};
  int p = bs;
  while (p < be) {
    const Char ch = decode(t, p, be);
    if (digit_value(ch.cp) < 0) { p += ch.len; continue; }
    const int d1 = skip_digits(t, p, be);
    int q = d1;
    if (q < be && t[q] == '.') ++q;
    q = skip_digits(t, q, be);
    int r = -1;
    if (q + 1 < be && t[q] == ' ') {
      const int c = q + 1;
      if (ignore_case) {
        if (c + 2 < be && ci_eq(t[c], 't') && ci_eq(t[c + 1], 'o') && t[c + 2] == ' ') r = c + 3;
        else if (c + 3 < be && ci_eq(t[c], 'a') && ci_eq(t[c + 1], 'n') && ci_eq(t[c + 2], 'd') && t[c + 3] == ' ') r = c + 4;
      } else {
        if (c + 2 < be && t[c] == 't' && t[c + 1] == 'o' && t[c + 2] == ' ') r = c + 3;
        else if (c + 3 < be && t[c] == 'a' && t[c + 1] == 'n' && t[c + 2] == 'd' && t[c + 3] == ' ') r = c + 4;
      }
    }
    if (r >= 0 && r < be) {
      int g = skip_digits(t, r, be);
      if (g > r) {
        if (g < be && t[g] == '.') ++g;
        g = skip_digits(t, g, be);
        last = Span{p, q, r, g};
        p = g;
        continue;
      }
    }
    p = d1;
  }
  return last;
}

// re.findall(r"\d+(?:\.\d+)?", text)[:2] (calc_difficulty.py:40-46).  Returns how many were found (0..2).
AS_HD int first_two_numbers(const uint8_t* t, int len, int* rng /*[4]*/) {
  int found = 0, p = 0;
  while (p < len && found < 2) {
    const Char ch = decode(t, p, len);
    if (digit_value(ch.cp) < 0) { p += ch.len; continue; }
    int q = skip_digits(t, p, len);
    if (q < len && t[q] == '.') {
      const int g = skip_digits(t, q + 1, len);
      if (g > q + 1) q = g;
    }
    rng[2 * found] = p;
    rng[2 * found + 1] = q;
    ++found;
    p = q;
  }
  return found;
}

struct Parsed {
  double s, e;
  uint8_t valid, format;
};

enum Mode {
  kTrainReward = 0,   // parse_timestamp_output (main.py:122-142) + format_reward (main.py:234-239)
  kEvalExtract = 1,   // extract_answer (evaluate.py:125-151)
  kEvalForce = 2,     // extract_answer, then extract_answer_force when it finds nothing (calc_difficulty.py:49-60)
};

// One text row.  Every lane of the policy runs this with identical control flow; lane 0 owns the sequential parts.
template <class W>
AS_HD Parsed scan_row(const uint8_t* t, int len, int mode, const W& w) {
  Parsed out{0.0, 0.0, 0, 0};
  int bs = 0, be = len;
  bool have_block = true;
  if (mode == kTrainReward) {
    // re.findall(r"<answer>(.*?)</answer>", s, DOTALL)[-1]: leftmost opening tag, first closing tag after it, resume
    // behind the closing tag.  Once an opening tag has no closing tag behind it, no later one has either.
    have_block = false;
    int pos = 0;
    for (;;) {
      const int i = find_pat(t, pos, len, "<answer>", 8, w);
      if (i < 0) break;
      const int j = find_pat(t, i + 8, len, "</answer>", 9, w);
      if (j < 0) break;
      bs = i + 8; be = j; have_block = true;
      pos = j + 9;
    }
  }
  if (have_block && w.lane() == 0) {
    // evaluate.py's fallback (the span inside the first single-line <answer> block) can never fire: a span inside a
    // substring is a span of the whole string, which was searched first
    const Span sp = last_span(t, bs, be, mode == kTrainReward);
    if (sp.s0 >= 0) {
      out.s = to_double(t, sp.s0, sp.s1);
      out.e = to_double(t, sp.e0, sp.e1);
      out.valid = 1;
    } else if (mode == kEvalForce) {
      int rng[4];
      if (first_two_numbers(t, len, rng) == 2) {
        out.s = to_double(t, rng[0], rng[1]);
        out.e = to_double(t, rng[2], rng[3]);
        out.valid = 1;
      }
    }
  }
  w.sync();
  if (mode != kTrainReward) return out;
  // format_reward: T = s.strip() must start with <think>, end with </answer>, and contain some "</think>" (at or after
  // byte 7) followed by optional whitespace and "<answer>" that ends before the final "</answer>" begins.  The lazy
  // quantifiers only decide WHICH split is found, not WHETHER one exists.
  int a = 0, b = len;
  while (a < len) {
    const Char ch = decode(t, a, len);
    if (!is_space(ch.cp)) break;
    a += ch.len;
  }
  while (b > a) {
    const int l = trailing_space_len(t, a, b);
    if (!l) break;
    b -= l;
  }
  if (b - a >= 7 + 8 + 8 + 9 && match_at(t, a, b, "<think>", 7) && match_at(t, b - 9, b, "</answer>", 9)) {
    const int limit = b - 9 - 8;  // last byte offset at which "<answer>" may start
    int pos = a + 7;
    for (;;) {
      const int i = find_pat(t, pos, b, "</think>", 8, w);
      if (i < 0) break;
      int r = i + 8;
      while (r < b) {
        const Char ch = decode(t, r, b);
        if (!is_space(ch.cp)) break;
        r += ch.len;
      }
      if (r <= limit && match_at(t, r, b, "<answer>", 8)) { out.format = 1; break; }
      pos = i + 1;
    }
  }
  return out;
}

}  // namespace ascan
