// cavlc_writer.h — slice_data() re-serialiser: macroblock records -> the original CAVLC bits.
//
// Decompress direction of the recompressor: the reference turns a DecodedMacroblock back into bits with
// OpenH264's ENCODER syntax writers (EncoderState::writeMbSynCavlc, codec/decoder/core/src/decode_slice.cpp:1442-1835 ->
// codec/encoder/core/src/svc_set_mb_syn_cavlc.cpp).  This is the host-side equivalent working on the hand-off records of
// include/lh264_b200.h (lh264_model_mb + the coefficient pool): H.264 7.3.4 / 7.3.5 / 9.2 in the writing direction, with the
// canonical choices every encoder makes (as many trailing ones as possible, shortest level escape).  The code tables are the
// (length, code) pairs of cavlc.cpp (Tables 9-5, 9-7 ... 9-10).
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <vector>
#include "../../include/lh264_b200.h"

namespace lh264host {

class BitWriter {
 public:
  void put (uint32_t v, int n) { for (int i = n - 1; i >= 0; i--) bit ((v >> i) & 1u); }
  void bit (uint32_t b) { if ((n_ & 7) == 0) buf_.push_back (0); if (b) buf_.back() |= (uint8_t) (0x80u >> (n_ & 7)); n_++; }
  void ue (uint32_t v) { const uint32_t x = v + 1; int len = 0; while ((x >> len) > 1) len++; put (0, len); put (x, len + 1); }
  void se (int32_t v) { ue (v > 0 ? (uint32_t) (2 * v - 1) : (uint32_t) (-2 * v)); }
  void te (int range, uint32_t v) { if (range > 1) ue (v); else bit (v ^ 1u); }
  size_t bits() const { return n_; }
  const std::vector<uint8_t>& bytes() const { return buf_; }
  void clear() { buf_.clear(); n_ = 0; }
 private:
  std::vector<uint8_t> buf_;
  size_t n_ = 0;
};

struct SliceWriteParams {
  int slice_type = 0;          // 0 P, 2 I
  int num_ref_idx_l0 = 1;
  int slice_qp = 26;
  bool transform_8x8_mode = false;
};

// One coded picture: keeps the per-block total_coeff and slice membership the nC context needs (9.2.1).
class PictureSyntaxWriter {
 public:
  PictureSyntaxWriter (int mb_w, int mb_h);
  void begin_slice (const SliceWriteParams& p);
  // Appends the syntax of one record (skip runs included) to bw.  Returns false on a record this writer cannot express
  // (I_PCM: the model carries no samples).
  bool write_mb (BitWriter& bw, const lh264_model_mb& r, const int16_t* coef_pool);
 private:
  int nC (int mb, int blk_raster, int plane /*0 luma, 1 Cb, 2 Cr*/) const;
  bool avail (int mb) const { return mb >= 0 && slice_of_[mb] == slice_no_; }
  int mb_w_, mb_h_, slice_no_ = 0, qp_ = 26;
  SliceWriteParams p_;
  std::vector<uint8_t> tc_;        // [mb][24] total_coeff: 16 luma (raster), 4 Cb, 4 Cr
  std::vector<int> slice_of_;
};

// residual_block_cavlc(): levels[] in scan order, max_coeff of them
void cavlc_write_residual_block (BitWriter& bw, int nC, int max_coeff, const int16_t* levels_in_scan_order);

} // namespace lh264host
