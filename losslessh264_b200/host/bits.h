// bits.h — RBSP bit reader (H.264 7.2 / 9.1 descriptors).  Replaces the reference's
// SBitStringAux + dec_golomb.h readers (codec/decoder/core/inc/dec_golomb.h) on the host side.
// The buffer handed in MUST be followed by at least 8 readable zero bytes (the NAL unescape
// step guarantees it), so peeks never need a bounds check.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string.h>

namespace lh264host {

class BitReader {
 public:
  BitReader() : p_ (nullptr), n_ (0), pos_ (0), payload_end_ (0) {}
  BitReader (const uint8_t* p, size_t nbytes) : p_ (p), n_ (nbytes * 8), pos_ (0) {
    // position of the rbsp_stop_one_bit = last 1 bit of the buffer
    size_t last = nbytes;
    while (last > 0 && p[last - 1] == 0) last--;
    if (last == 0) payload_end_ = 0;
    else {
      const uint8_t b = p[last - 1];
      int tz = 0;
      while (! ((b >> tz) & 1)) tz++;
      payload_end_ = (last - 1) * 8 + (7 - tz);   // bit index of the stop bit
    }
  }
  size_t pos() const { return pos_; }
  size_t size_bits() const { return n_; }
  void seek (size_t bitpos) { pos_ = bitpos; }
  bool overrun() const { return pos_ > n_; }
  const uint8_t* data() const { return p_; }

  inline uint32_t peek (int n) const {   // 1..32 bits, MSB first; past the end of the buffer (damaged streams) reads are zeros
    if (pos_ > n_) return 0;             // (pos_ <= n_ keeps the 8-byte load inside the zero padding the caller guarantees)
    uint64_t v;
    memcpy (&v, p_ + (pos_ >> 3), 8);
    v = __builtin_bswap64 (v) << (pos_ & 7);
    return (uint32_t) (v >> (64 - n));
  }
  inline void skip (int n) { pos_ += n; }
  inline uint32_t u (int n) { if (n == 0) return 0; uint32_t v = peek (n); pos_ += n; return v; }
  inline uint32_t u1() { if (pos_ > n_) { pos_++; return 0; } uint32_t v = (p_[pos_ >> 3] >> (7 - (pos_ & 7))) & 1; pos_++; return v; }
  inline uint32_t ue() {
    const uint32_t w = peek (32);
    if (w == 0) {                       // >= 32 leading zeros: malformed for every syntax element we read
      pos_ += 32;
      return 0xffffffffu;
    }
    const int zeros = __builtin_clz (w);
    if (zeros <= 15) {                  // whole code word inside the 32-bit window
      pos_ += 2 * zeros + 1;
      return (w >> (31 - 2 * zeros)) - 1;
    }
    pos_ += zeros + 1;
    return (uint32_t) ((1ull << zeros) - 1 + u (zeros));
  }
  inline int32_t se() {
    const uint32_t k = ue();
    return (k & 1) ? (int32_t) ((k + 1) >> 1) : - (int32_t) (k >> 1);
  }
  inline uint32_t te (int range) {      // range = largest value; range == 1 -> inverted single bit
    if (range > 1) return ue();
    return u1() ^ 1;
  }
  bool byte_aligned() const { return (pos_ & 7) == 0; }
  void align() { pos_ = (pos_ + 7) & ~ (size_t)7; }
  // 7.2 more_rbsp_data(): true unless only rbsp_trailing_bits remain
  bool more_rbsp_data() const { return pos_ < payload_end_; }
  size_t payload_end() const { return payload_end_; }
 private:
  const uint8_t* p_;
  size_t n_, pos_, payload_end_;
};

} // namespace lh264host
