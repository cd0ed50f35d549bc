// sinks_gpu.cpp — GpuSink, see sinks.h.
#include "sinks.h"
#include <string.h>

namespace lh264host {

GpuSink::GpuSink (lh264_ctx* shared, int device) : ctx_ (shared), own_ctx_ (false) {
  if (!ctx_) { if (lh264_ctx_create (&ctx_, device, nullptr) == LH264_OK) own_ctx_ = true; else ctx_ = nullptr; }
}
GpuSink::~GpuSink() {
  if (ctx_ && stream_ >= 0) lh264_stream_close (ctx_, stream_);
  if (ctx_ && own_ctx_) lh264_ctx_destroy (ctx_);
}
const char* GpuSink::error() const { return lh264_last_error(); }
int GpuSink::configure (int mb_w, int mb_h, int n_slots) {
  if (!ctx_) return LH264_E_NO_DEVICE;
  if (stream_ >= 0) lh264_stream_close (ctx_, stream_);
  stream_ = lh264_stream_open (ctx_, mb_w, mb_h, n_slots);
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
int GpuSink::read (const DecodedPictureInfo& info, uint8_t* y, int sy, uint8_t* u, uint8_t* v, int sc) {
  if (!ctx_ || stream_ < 0) return LH264_E_NO_DEVICE;
  return lh264_picture_read (ctx_, stream_, info.slot, info.crop_x, info.crop_y, info.width, info.height, y, sy, u, v, sc, 1);
}

} // namespace lh264host
