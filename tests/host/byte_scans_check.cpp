// byte_scans_check.cpp — test driver (tests/test_byte_scans.py): the memchr-based start-code scans and NAL unescape of the host
// parser (h264_syntax.cpp: split_annexb, unescape_nal; pip_container.cpp: PipBoilerplate::next_unit) against byte-by-byte
// restatements of B.1 / 7.4.1, on random buffers dense in zeros, 00 00 01, 00 00 00 01 and 00 00 03 patterns.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "h264_syntax.h"
#include "pip_container.h"

using namespace lh264host;

static uint64_t rng_state = 0x9e3779b97f4a7c15ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t) (rng_state >> 16); }

static void naive_split (const uint8_t* d, size_t len, std::vector<NalSpan>& out) {
  out.clear();
  auto find_sc = [&] (size_t from, size_t& scl) -> size_t {
    for (size_t k = from; k + 3 <= len; k++)
      if (d[k] == 0 && d[k + 1] == 0) {
        if (d[k + 2] == 1) { scl = 3; return k; }
        if (k + 4 <= len && d[k + 2] == 0 && d[k + 3] == 1) { scl = 4; return k; }
      }
    scl = 0;
    return len;
  };
  size_t scl = 0, i = find_sc (0, scl);
  while (i < len) {
    const size_t start = i + scl;
    size_t nscl = 0;
    const size_t next = find_sc (start, nscl);
    size_t end = next;
    while (end > start && d[end - 1] == 0) end--;
    if (end > start) out.push_back (NalSpan{d + start, end - start});
    i = next; scl = nscl;
  }
}

static size_t naive_unescape (const uint8_t* p, size_t len, std::vector<uint8_t>& out) {
  out.clear();
  int zeros = 0;
  for (size_t i = 0; i < len; i++) {
    const uint8_t b = p[i];
    if (zeros >= 2 && b == 3) { zeros = 0; continue; }
    out.push_back (b);
    zeros = b == 0 ? zeros + 1 : 0;
  }
  return out.size();
}

static size_t naive_next_unit (const uint8_t* b, size_t pos, size_t file_size) {
  size_t i;
  for (i = 0; i < file_size; i++) {
    const uint8_t* q = b + pos + i;
    if (i > 0 && q[0] == 0 && q[1] == 0 && ((q[2] == 0 && q[3] == 1) || q[2] == 1)) break;
  }
  return i;
}

int main (int argc, char** argv) {
  const int cases = argc > 1 ? atoi (argv[1]) : 2000;
  std::vector<uint8_t> reuse;                              // unescape_nal into a REUSED vector (stale bytes behind the result)
  for (int c = 0; c < cases; c++) {
    const size_t len = rnd() % 700;
    const int zero_pct = (int) (rnd() % 60);               // how dense the zeros are
    std::vector<uint8_t> buf (len + 4 + 16, 0xAA);
    for (size_t i = 0; i < len; i++) {
      const uint32_t r = rnd() % 100;
      buf[i] = r < (uint32_t)zero_pct ? 0 : (r < (uint32_t)zero_pct + 12 ? (uint8_t) (1 + rnd() % 3) : (uint8_t)rnd());
    }
    static const uint8_t sentinel[4] = {0, 0, 0, 1};
    memcpy (buf.data() + len, sentinel, 4);                // what the runtime appends behind the last byte
    // split_annexb
    std::vector<NalSpan> a, b;
    split_annexb (buf.data(), len, a);
    naive_split (buf.data(), len, b);
    if (a.size() != b.size()) { fprintf (stderr, "case %d: split_annexb finds %zu units, the plain walk %zu\n", c, a.size(), b.size()); return 1; }
    for (size_t i = 0; i < a.size(); i++) if (a[i].p != b[i].p || a[i].len != b[i].len) { fprintf (stderr, "case %d: unit %zu differs\n", c, i); return 1; }
    // unescape_nal (whole buffer and every unit)
    std::vector<uint8_t> want;
    const size_t n = unescape_nal (buf.data(), len, reuse), m = naive_unescape (buf.data(), len, want);
    if (n != m || memcmp (reuse.data(), want.data(), m) != 0) { fprintf (stderr, "case %d: unescape_nal differs (%zu vs %zu bytes)\n", c, n, m); return 1; }
    for (int k = 0; k < 8; k++) if (reuse[n + k] != 0) { fprintf (stderr, "case %d: no zero padding behind the payload\n", c); return 1; }
    // next_unit from every position that could start a unit
    for (size_t pos = 0; pos < len; pos += 1 + rnd() % 7) {
      const size_t x = PipBoilerplate::next_unit (buf.data(), pos, len), y = naive_next_unit (buf.data(), pos, len);
      if (x != y) { fprintf (stderr, "case %d: next_unit at %zu gives %zu, the plain walk %zu\n", c, pos, x, y); return 1; }
    }
  }
  printf ("ok %d cases\n", cases);
  return 0;
}
