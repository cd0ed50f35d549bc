// cavlc.cpp — see cavlc.h.
#include "cavlc.h"
#include "cavlc_tables.h"
#include <string.h>
#include <vector>
#include <mutex>

namespace lh264host {

const uint8_t kZigzag4x4[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};
const uint8_t kZigzag8x8[64] = {
  0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
  35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63
};

// the (length, code) tables, for the writing direction (cavlc_writer.cpp)
void cavlc_tables (const uint8_t (**tok_len)[4][17], const uint8_t (**tok_code)[4][17], const uint8_t (**tokc_len)[5], const uint8_t (**tokc_code)[5],
                   const uint8_t (**tz_len)[16], const uint8_t (**tz_code)[16], const uint8_t (**tzc_len)[4], const uint8_t (**tzc_code)[4],
                   const uint8_t (**run_len)[15], const uint8_t (**run_code)[15]) {
  *tok_len = kTokLen; *tok_code = kTokCode; *tokc_len = kTokLenC; *tokc_code = kTokCodeC; *tz_len = kTzLen; *tz_code = kTzCode;
  *tzc_len = kTzLenC; *tzc_code = kTzCodeC; *run_len = kRunLen; *run_code = kRunCode;
}

// ---- generic prefix-code LUT ------------------------------------------------------------------------------------
struct Lut {
  // two levels: the top 8 bits resolve every code of up to 8 bits (the bulk of what real streams hold) from a 512-byte table that
  // stays in L1; longer codes go through the full-width table.  Entry = (symbol << 8) | length, 0 = not a code.
  int bits = 0;                         // index width of the full table (<= 16)
  uint16_t l1[256];
  std::vector<uint16_t> full;
  void build (int nbits, const uint8_t* lens, const uint8_t* codes, int n) {
    bits = nbits;
    memset (l1, 0, sizeof (l1));
    full.assign ((size_t)1 << nbits, 0);
    for (int s = 0; s < n; s++) {
      const int l = lens[s];
      if (!l) continue;
      const uint16_t e = (uint16_t) ((s << 8) | l);
      const uint32_t first = (uint32_t)codes[s] << (nbits - l), cnt = 1u << (nbits - l);
      for (uint32_t k = 0; k < cnt; k++) full[first + k] = e;
      if (l <= 8) { const uint32_t f1 = (uint32_t)codes[s] << (8 - l), c1 = 1u << (8 - l); for (uint32_t k = 0; k < c1; k++) l1[f1 + k] = e; }
    }
  }
  inline int get (BitReader& br) const {
    const uint32_t w = br.peek (16);
    uint32_t e = l1[w >> 8];
    if (!e) {
      e = full[w >> (16 - bits)];
      if (!e) return -1;
    }
    br.skip ((int) (e & 0xffu));
    return (int) (e >> 8);
  }
};

static Lut g_tok[3], g_tokc, g_tz[15], g_tzc[3], g_run[7];
static std::once_flag g_once;

static void do_init() {
  for (int t = 0; t < 3; t++) {
    uint8_t lens[68], codes[68];
    for (int t1 = 0; t1 < 4; t1++) for (int tc = 0; tc < 17; tc++) { lens[tc * 4 + t1] = kTokLen[t][t1][tc]; codes[tc * 4 + t1] = kTokCode[t][t1][tc]; }
    g_tok[t].build (16, lens, codes, 68);
  }
  {
    uint8_t lens[20], codes[20];
    for (int t1 = 0; t1 < 4; t1++) for (int tc = 0; tc < 5; tc++) { lens[tc * 4 + t1] = kTokLenC[t1][tc]; codes[tc * 4 + t1] = kTokCodeC[t1][tc]; }
    g_tokc.build (8, lens, codes, 20);
  }
  for (int i = 0; i < 15; i++) g_tz[i].build (9, kTzLen[i], kTzCode[i], 16 - i);
  for (int i = 0; i < 3; i++) g_tzc[i].build (3, kTzLenC[i], kTzCodeC[i], 4 - i);
  for (int i = 0; i < 7; i++) g_run[i].build (11, kRunLen[i], kRunCode[i], i < 6 ? i + 2 : 15);
}
void cavlc_init() { std::call_once (g_once, do_init); }

static bool kraft_ok (const uint8_t* lens, const uint8_t* codes, int n, bool complete) {
  // prefix-free and (optionally) complete
  double sum = 0;
  for (int a = 0; a < n; a++) {
    if (!lens[a]) continue;
    sum += 1.0 / (double) (1u << lens[a]);
    for (int b = 0; b < n; b++) {
      if (a == b || !lens[b] || lens[b] < lens[a]) continue;
      if ((uint32_t) (codes[b] >> (lens[b] - lens[a])) == codes[a]) return false;
    }
  }
  return complete ? (sum > 0.999999 && sum < 1.000001) : sum <= 1.000001;
}
bool cavlc_selfcheck() {
  bool ok = true;
  for (int t = 0; t < 3; t++) {
    uint8_t lens[68], codes[68];
    for (int t1 = 0; t1 < 4; t1++) for (int tc = 0; tc < 17; tc++) { lens[tc * 4 + t1] = kTokLen[t][t1][tc]; codes[tc * 4 + t1] = kTokCode[t][t1][tc]; }
    ok &= kraft_ok (lens, codes, 68, false);
  }
  { uint8_t lens[20], codes[20];
    for (int t1 = 0; t1 < 4; t1++) for (int tc = 0; tc < 5; tc++) { lens[tc * 4 + t1] = kTokLenC[t1][tc]; codes[tc * 4 + t1] = kTokCodeC[t1][tc]; }
    ok &= kraft_ok (lens, codes, 20, false); }
  for (int i = 0; i < 15; i++) ok &= kraft_ok (kTzLen[i], kTzCode[i], 16 - i, i != 0 ? true : false);
  for (int i = 0; i < 3; i++) ok &= kraft_ok (kTzLenC[i], kTzCodeC[i], 4 - i, true);
  for (int i = 0; i < 7; i++) ok &= kraft_ok (kRunLen[i], kRunCode[i], i < 6 ? i + 2 : 15, i < 6);
  return ok;
}

#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__) && !defined(LH264_NO_CLONES)
__attribute__ ((target_clones ("arch=x86-64-v3", "default")))       // see decoder.cpp LH264_HOT
#endif
int cavlc_residual_block (BitReader& br, int nC, int max_coeff, const uint8_t* scan, int scan_step, int scan_off, int16_t* out) {
  int total, t1;
  if (nC < 0) {
    const int s = g_tokc.get (br);
    if (s < 0) return -1;
    total = s >> 2; t1 = s & 3;
  } else if (nC >= 8) {                      // 6-bit FLC: xxxxyy, 000011 = no coefficients
    const uint32_t v = br.u (6);
    if (v == 3) { total = 0; t1 = 0; }
    else { total = (v >> 2) + 1; t1 = v & 3; }
  } else {
    const int s = g_tok[nC < 2 ? 0 : (nC < 4 ? 1 : 2)].get (br);
    if (s < 0) return -1;
    total = s >> 2; t1 = s & 3;
  }
  if (total == 0) return 0;
  if (total > max_coeff || t1 > total) return -1;
  int level[16];
  // 9.2.2: trailing ones then levels, highest frequency first
  for (int i = 0; i < t1; i++) level[i] = br.u1() ? -1 : 1;
  int suffix_len = (total > 10 && t1 < 3) ? 1 : 0;
  for (int i = t1; i < total; i++) {
    const uint32_t w = br.peek (32);
    if (w == 0) return -1;
    const int prefix = __builtin_clz (w);
    br.skip (prefix + 1);
    int code = (prefix < 15 ? prefix : 15) << suffix_len;
    int suffix_size = suffix_len;
    if (prefix == 14 && suffix_len == 0) suffix_size = 4;
    else if (prefix >= 15) suffix_size = prefix - 3;
    if (suffix_size > 0) code += (int)br.u (suffix_size);
    if (prefix >= 15 && suffix_len == 0) code += 15;
    if (prefix >= 16) code += (1 << (prefix - 3)) - 4096;
    if (i == t1 && t1 < 3) code += 2;
    const int lv = (code & 1) ? (-code - 1) >> 1 : (code + 2) >> 1;
    level[i] = lv;
    if (suffix_len == 0) suffix_len = 1;
    if ((lv < 0 ? -lv : lv) > (3 << (suffix_len - 1)) && suffix_len < 6) suffix_len++;
  }
  int zeros_left = 0;
  if (total < max_coeff) {
    int tz;
    if (nC < 0) tz = g_tzc[total - 1].get (br);
    else tz = g_tz[total - 1].get (br);
    if (tz < 0) return -1;
    zeros_left = tz;
  }
  if (zeros_left + total > max_coeff) return -1;
  // place coefficients: level[0] is the highest-frequency one
  int pos = zeros_left + total - 1;
  for (int i = 0; i < total; i++) {
    out[scan[scan_off + pos * scan_step]] = (int16_t)level[i];
    if (i + 1 < total && zeros_left > 0) {
      const int run = g_run[(zeros_left > 7 ? 7 : zeros_left) - 1].get (br);
      if (run < 0 || run > zeros_left) return -1;
      zeros_left -= run;
      pos -= run;
    }
    pos--;
  }
  return total;
}

} // namespace lh264host
