// sinks.h — PictureSink implementations: the GPU sink (the product path: C ABI of
// include/lh264_b200.h, loaded from liblh264_b200.so) and a file sink that writes .lhfb records
// (tests / tools; no reconstruction happens there).
#pragma once
#include <stdio.h>
#include <string>
#include <vector>
#include "decoder.h"

namespace lh264host {

// Writes one .lhfb slice record per slice (format: oracle/ref_dump_hook.cpp header comment).
class LhfbSink : public PictureSink {
 public:
  explicit LhfbSink (const char* path);
  ~LhfbSink() override;
  int configure (int mb_w, int mb_h, int n_slots) override;
  int submit (const lh264_picture_desc& pic, const DecodedPictureInfo& info) override;
  // optional second file: the recompression-model hand-off, one 'LHMD' record per picture:
  //   u32 magic 'LHMD'; i32 mb_w, mb_h, n_slices, n_recs, n_coef_blocks; lh264_model_slice[n_slices];
  //   lh264_model_mb[n_recs]; int16[n_coef_blocks][16]
  void set_model_path (const char* path);
  int submit_model (const lh264_model_picture& pic, const DecodedPictureInfo& info) override;
  bool ok() const { return f_ != nullptr; }
 private:
  FILE* f_;
  FILE* fm_ = nullptr;
  int mb_w_ = 0, mb_h_ = 0;
};

// Reconstructs on the GPU through the C ABI.  One stream of one lh264_ctx.
class GpuSink : public PictureSink {
 public:
  // ctx may be shared between sinks (one per decoder instance); NULL = create a private one.
  explicit GpuSink (lh264_ctx* shared_ctx = nullptr, int device = 0);
  ~GpuSink() override;
  int configure (int mb_w, int mb_h, int n_slots) override;
  int submit (const lh264_picture_desc& pic, const DecodedPictureInfo& info) override;
  int fill_slot (int slot, int value) override;
  // recompression model of the picture on the GPU (lh264_model_submit); the model stream is opened on first use
  int submit_model (const lh264_model_picture& pic, const DecodedPictureInfo& info) override;
  // ends the model's tag streams; bytes of one `.pip.<tag>` side stream (empty = the tag has no file)
  int model_finish();
  int sync();                              // everything submitted so far is done (deferred coding included)
  int model_tag_bytes_from (int tag, size_t offset, std::vector<uint8_t>& out);   // the bytes of a tag stream from `offset` on
  int model_tag_bytes (int tag, std::vector<uint8_t>& out);
  // cropped I420 of a decoded picture into host memory (synchronous)
  int read (const DecodedPictureInfo& info, uint8_t* y, int stride_y, uint8_t* u, uint8_t* v, int stride_c);
  lh264_ctx* ctx() const { return ctx_; }
  int stream() const { return stream_; }
  const char* error() const;
 private:
  lh264_ctx* ctx_;
  bool own_ctx_;
  int stream_ = -1;
  int model_ = -1, model_mbs_ = 0;
  int w_ = 0, h_ = 0;
};

} // namespace lh264host
