// cabac.h — CABAC arithmetic decoding engine and context set (H.264 9.3.1 initialisation, 9.3.3.2 decoding).  Host side;
// replaces the reference's cabac_decoder.cpp (InitCabacDecEngineFromBS, DecodeBinCabac, DecodeBypassCabac, DecodeTerminateCabac,
// codec/decoder/core/src/cabac_decoder.cpp:38-320).  The arithmetic is the specification's (9-bit codIRange / codIOffset); the
// bitstream is pulled in 32 bits at a time and the exact read position is handed back to the RBSP reader whenever a terminate
// bin says 1 (end_of_slice_flag, I_PCM).
#pragma once
#include <stdint.h>
#include "bits.h"
#include "cabac_tables.h"

namespace lh264host {

class CabacDecoder {
 public:
  // 9.3.1.1: context variables from (m, n) and SliceQPY; cls = 0 for I slices, 1 + cabac_init_idc otherwise
  void init_contexts (int cls, int slice_qp) {
    const int qp = slice_qp < 0 ? 0 : (slice_qp > 51 ? 51 : slice_qp);
    for (int i = 0; i < 460; i++) {
      const int v = kCabacInitMN[cls][i], m = v >> 8, n = (int8_t) (v & 0xff);
      int pre = ((m * qp) >> 4) + n;
      pre = pre < 1 ? 1 : (pre > 126 ? 126 : pre);
      state_[i] = pre <= 63 ? (uint8_t) ((63 - pre) << 1) : (uint8_t) (((pre - 64) << 1) | 1);
    }
  }
  // 9.3.1.2: the reader must be byte aligned (cabac_alignment_one_bit / pcm samples consumed).
  // The engine keeps codIOffset followed by `avail_` prefetched bitstream bits in one 64-bit word (val_ = codIOffset << avail_ | bits),
  // so a renormalisation is a subtraction from avail_ and memory is touched once per 32 bits; comparisons scale codIRange
  // instead.  `fill_` is the bit position up to which the stream has been pulled in: the specification's read position is
  // fill_ - avail_, and it is written back to the reader whenever a terminate bin says 1 (end_of_slice_flag, I_PCM).
  void start (BitReader* br) {
    br_ = br; p_ = br->data(); n_bits_ = br->size_bits(); fill_ = br->pos();
    range_ = 510; val_ = 0; avail_ = -9;
    refill();                                          // 32 bits in: 9 of them are codIOffset
  }

  inline int decision (int ctx) {                     // 9.3.3.2.1 DecodeDecision
    uint8_t& s = state_[ctx];
    const uint32_t idx = s >> 1;
    uint32_t mps = s & 1u;
    const uint32_t lps = kCabacRangeLps[(range_ >> 6) & 3][idx];
    range_ -= lps;
    const uint64_t scaled = (uint64_t)range_ << avail_;
    int bin;
    if (val_ >= scaled) {
      bin = (int) (mps ^ 1u);
      val_ -= scaled;
      range_ = lps;
      if (idx == 0) mps ^= 1u;
      s = (uint8_t) ((kCabacNextLps[idx] << 1) | mps);
    } else {
      bin = (int)mps;
      s = (uint8_t) (((idx < 62 ? idx + 1 : idx) << 1) | mps);
    }
    if (range_ < 256) {                               // RenormD, all shifts at once
      const int sh = __builtin_clz (range_) - 23;
      range_ <<= sh;
      if (avail_ < sh) refill();
      avail_ -= sh;
    }
    return bin;
  }
  inline int bypass() {                               // 9.3.3.2.3
    if (avail_ < 1) refill();
    avail_--;
    const uint64_t scaled = (uint64_t)range_ << avail_;
    if (val_ >= scaled) { val_ -= scaled; return 1; }
    return 0;
  }
  inline int terminate() {                            // 9.3.3.2.2.3; returns 1 without renormalising: the reader then stands right
    range_ -= 2;                                      // behind the last bit the encoder's flush wrote
    if (val_ >= ((uint64_t)range_ << avail_)) { br_->seek (fill_ - (size_t)avail_); return 1; }
    if (range_ < 256) { range_ <<= 1; if (avail_ < 1) refill(); avail_--; }
    return 0;
  }
  bool overrun() const { return fill_ - (size_t)avail_ > n_bits_; }

 private:
  inline void refill() {                              // 32 more bits behind val_ (zeros past the end of the RBSP)
    uint64_t w = 0;
    if (fill_ <= n_bits_) {                           // (the reader's buffer is followed by 8 readable zero bytes)
      memcpy (&w, p_ + (fill_ >> 3), 8);
      w = (__builtin_bswap64 (w) << (fill_ & 7)) >> 32;
    }
    val_ = (val_ << 32) | w;
    fill_ += 32; avail_ += 32;
  }
  BitReader* br_ = nullptr;
  const uint8_t* p_ = nullptr;
  size_t n_bits_ = 0, fill_ = 0;
  uint64_t val_ = 0;
  int avail_ = 0;
  uint32_t range_ = 510;
  uint8_t state_[460];                                // (pStateIdx << 1) | valMPS
};

} // namespace lh264host
