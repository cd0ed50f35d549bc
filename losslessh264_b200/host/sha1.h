// sha1.h — incremental SHA-1 (FIPS 180-1) for hashing decoded pictures on the host; the reference's tests pin decoded YUV by
// SHA1 (test/api/decoder_test.cpp:6-12, :63-70).
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <string>

namespace lh264host {

class Sha1 {
 public:
  Sha1() { reset(); }
  void reset() { h_[0] = 0x67452301u; h_[1] = 0xEFCDAB89u; h_[2] = 0x98BADCFEu; h_[3] = 0x10325476u; h_[4] = 0xC3D2E1F0u; len_ = 0; fill_ = 0; }
  void update (const uint8_t* p, size_t n) {
    len_ += n;
    if (fill_) {
      const size_t k = n < 64 - fill_ ? n : 64 - fill_;
      memcpy (buf_ + fill_, p, k); fill_ += k; p += k; n -= k;
      if (fill_ == 64) { block (buf_); fill_ = 0; }
    }
    while (n >= 64) { block (p); p += 64; n -= 64; }
    if (n) { memcpy (buf_, p, n); fill_ = n; }
  }
  std::string hex() {
    uint64_t bits = len_ * 8;
    uint8_t pad[72];
    size_t k = 0;
    pad[k++] = 0x80;
    while ((fill_ + k) % 64 != 56) pad[k++] = 0;
    for (int i = 7; i >= 0; i--) pad[k++] = (uint8_t) (bits >> (8 * i));
    update (pad, k);
    static const char* d = "0123456789abcdef";
    std::string s;
    for (int i = 0; i < 5; i++) for (int j = 28; j >= 0; j -= 4) s += d[(h_[i] >> j) & 15];
    return s;
  }
 private:
  static uint32_t rol (uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
  void block (const uint8_t* p) {
    uint32_t w[80];
    for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
    for (int i = 16; i < 80; i++) w[i] = rol (w[i - 3] ^ w[i - 8] ^ w[i - 14] ^ w[i - 16], 1);
    uint32_t a = h_[0], b = h_[1], c = h_[2], d = h_[3], e = h_[4];
    for (int i = 0; i < 80; i++) {
      uint32_t f, k;
      if (i < 20) { f = (b & c) | (~b & d); k = 0x5A827999u; }
      else if (i < 40) { f = b ^ c ^ d; k = 0x6ED9EBA1u; }
      else if (i < 60) { f = (b & c) | (b & d) | (c & d); k = 0x8F1BBCDCu; }
      else { f = b ^ c ^ d; k = 0xCA62C1D6u; }
      const uint32_t t = rol (a, 5) + f + e + k + w[i];
      e = d; d = c; c = rol (b, 30); b = a; a = t;
    }
    h_[0] += a; h_[1] += b; h_[2] += c; h_[3] += d; h_[4] += e;
  }
  uint32_t h_[5];
  uint64_t len_;
  uint8_t buf_[64];
  size_t fill_;
};

} // namespace lh264host
