// cabac.h — CABAC arithmetic decoding engine and context set (H.264 9.3.1 initialisation, 9.3.3.2 decoding).  Host side;
// replaces the reference's cabac_decoder.cpp (InitCabacDecEngineFromBS, DecodeBinCabac, DecodeBypassCabac, DecodeTerminateCabac,
// codec/decoder/core/src/cabac_decoder.cpp:38-320).  The engine is the specification's: a 9-bit codIRange / codIOffset pair
// that pulls bits from the RBSP reader one renormalisation at a time, so the reader's position is exact whenever a terminate
// bin says 1 (end_of_slice_flag, I_PCM) — no look-ahead has to be handed back.
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
  // 9.3.1.2: the reader must be byte aligned (cabac_alignment_one_bit / pcm samples consumed)
  void start (BitReader* br) { br_ = br; range_ = 510; offset_ = br->u (9); }

  inline int decision (int ctx) {                     // 9.3.3.2.1 DecodeDecision
    uint8_t& s = state_[ctx];
    const uint32_t idx = s >> 1;
    uint32_t mps = s & 1u;
    const uint32_t lps = kCabacRangeLps[(range_ >> 6) & 3][idx];
    range_ -= lps;
    int bin;
    if (offset_ >= range_) {
      bin = (int) (mps ^ 1u);
      offset_ -= range_;
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
      offset_ = (offset_ << sh) | br_->u (sh);
    }
    return bin;
  }
  inline int bypass() {                               // 9.3.3.2.3
    offset_ = (offset_ << 1) | br_->u1();
    if (offset_ >= range_) { offset_ -= range_; return 1; }
    return 0;
  }
  inline int terminate() {                            // 9.3.3.2.2.3; returns 1 without renormalising: the reader then stands right
    range_ -= 2;                                      // behind the last bit the encoder's flush wrote
    if (offset_ >= range_) return 1;
    if (range_ < 256) { range_ <<= 1; offset_ = (offset_ << 1) | br_->u1(); }
    return 0;
  }
  bool overrun() const { return br_->overrun(); }

 private:
  BitReader* br_ = nullptr;
  uint32_t range_ = 510, offset_ = 0;
  uint8_t state_[460];                                // (pStateIdx << 1) | valMPS
};

} // namespace lh264host
