// sinks_gpu.cpp — GpuSink, see sinks.h.
#include "sinks.h"
#include <string.h>

namespace lh264host {

GpuSink::GpuSink (lh264_ctx* shared, int device) : ctx_ (shared), own_ctx_ (false) {
  if (!ctx_) { if (lh264_ctx_create (&ctx_, device, nullptr) == LH264_OK) own_ctx_ = true; else ctx_ = nullptr; }
}
GpuSink::~GpuSink() {
  if (ctx_ && model_ >= 0) lh264_model_close (ctx_, model_);
  if (ctx_ && stream_ >= 0) lh264_stream_close (ctx_, stream_);
  if (ctx_ && own_ctx_) lh264_ctx_destroy (ctx_);
}
const char* GpuSink::error() const { return lh264_last_error(); }
int GpuSink::configure (int mb_w, int mb_h, int n_slots) {
  if (!ctx_) return LH264_E_NO_DEVICE;
  if (stream_ >= 0) lh264_stream_close (ctx_, stream_);
  stream_ = lh264_stream_open (ctx_, mb_w, mb_h, n_slots);
  w_ = mb_w * 16; h_ = mb_h * 16;
  return stream_ < 0 ? stream_ : 0;
}
int GpuSink::submit (const lh264_picture_desc& pic, const DecodedPictureInfo&) {
  if (!ctx_ || stream_ < 0) return LH264_E_NO_DEVICE;
  lh264_picture_desc d = pic;
  d.stream = stream_;
  int rc = lh264_recon_submit (ctx_, &d, 1);
  if (rc) return rc;
  return lh264_sync (ctx_);          // the decoder reuses its record buffers for the next picture
}
int GpuSink::fill_slot (int slot, int value) {
  if (!ctx_ || stream_ < 0 || w_ <= 0) return LH264_E_NO_DEVICE;
  std::vector<uint8_t> y ((size_t)w_ * h_, (uint8_t)value), c ((size_t)w_ * h_ / 4, (uint8_t)value);
  return lh264_picture_write (ctx_, stream_, slot, y.data(), w_, c.data(), c.data(), w_ / 2);
}
int GpuSink::submit_model (const lh264_model_picture& pic, const DecodedPictureInfo&) {
  if (!ctx_) return LH264_E_NO_DEVICE;
  const int mbs = pic.mb_w * pic.mb_h;
  if (model_ >= 0 && mbs > model_mbs_) return LH264_E_UNSUPPORTED;   // beyond level 5.2: not a stream the reference decodes either
  if (model_ < 0) {
    // The model's priors live across a change of resolution (the reference only drops its FreqImage then), so the stream is
    // opened once, for the largest picture an H.264 level allows (level 5.2: 36 864 macroblocks; this sizes two frames of 32-byte
    // context records, 2.4 MB) - a sequence that switches to a bigger picture (Error_I_P.264) keeps coding into the same state.
    model_mbs_ = mbs > 36864 ? mbs : 36864;
    model_ = lh264_model_open (ctx_, model_mbs_, /*with_8x8*/1);
    if (model_ < 0) return model_;
  }
  lh264_model_picture p = pic;
  p.model_stream = model_;
  return lh264_model_submit (ctx_, &p, 1);
}
int GpuSink::sync() { return ctx_ ? lh264_sync (ctx_) : LH264_OK; }
int GpuSink::model_tag_bytes_from (int tag, size_t offset, std::vector<uint8_t>& out) {
  out.clear();
  if (!ctx_ || model_ < 0) return LH264_OK;
  const long long n = lh264_model_read (ctx_, model_, tag, nullptr, 0, 0);
  if (n < 0) return (int)n;
  if ((size_t)n <= offset) return LH264_OK;
  out.resize ((size_t)n - offset);
  const long long m = lh264_model_read (ctx_, model_, tag, out.data(), offset, out.size());
  return m < 0 ? (int)m : LH264_OK;
}
int GpuSink::model_finish() { return ctx_ && model_ >= 0 ? lh264_model_finish (ctx_, model_) : LH264_OK; }
int GpuSink::model_tag_bytes (int tag, std::vector<uint8_t>& out) {
  out.clear();
  if (!ctx_ || model_ < 0) return LH264_OK;
  const long long n = lh264_model_read (ctx_, model_, tag, nullptr, 0, 0);
  if (n < 0) return (int)n;
  out.resize ((size_t)n);
  if (n > 0) { const long long m = lh264_model_read (ctx_, model_, tag, out.data(), 0, (size_t)n); if (m < 0) return (int)m; }
  return LH264_OK;
}
int GpuSink::read (const DecodedPictureInfo& info, uint8_t* y, int sy, uint8_t* u, uint8_t* v, int sc) {
  if (!ctx_ || stream_ < 0) return LH264_E_NO_DEVICE;
  return lh264_picture_read (ctx_, stream_, info.slot, info.crop_x, info.crop_y, info.width, info.height, y, sy, u, v, sc, 1);
}

} // namespace lh264host
