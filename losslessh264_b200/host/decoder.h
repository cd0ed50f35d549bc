// decoder.h — host decoder core: NAL -> parameter sets / slices -> per-picture hand-off records.
//
// This is the HOST side of the hot path (it stays on the CPU by design; one instance per
// stream, instances are independent so N host threads can each own one).  It plays the role of
// the reference's decoder framework + MB parse:
//   WelsDecodeBs            codec/decoder/core/src/decoder.cpp:656-859
//   DecodeCurrentAccessUnit codec/decoder/core/src/decoder_core.cpp:2118-2417
//   WelsDecodeSlice / WelsActualDecodeMbCavlc{I,P}Slice  decode_slice.cpp:2959-3171, :3173-3984
//   parse_mb_syn_cavlc.cpp, mv_pred.cpp (PredMv :46, PredPSkipMvFromNeighbor :197, 16x8 :244, 8x16 :267)
//   manage_dec_ref.cpp (WelsInitRefList :112, WelsReorderRefList :174, WelsMarkAsRef :254)
// and produces, instead of calling WelsTargetSliceConstruction per slice, one
// lh264_picture_desc per coded picture that a PictureSink consumes (GPU sink = the product;
// dump sink = tests).  Written from the H.264 specification, not from the reference sources.
#pragma once
#include <stdint.h>
#include <vector>
#include <memory>
#include "../../include/lh264_b200.h"
#include "h264_syntax.h"

namespace lh264host {

struct DecodedPictureInfo {
  int mb_w, mb_h;
  int crop_x, crop_y, width, height;    // cropped output window (luma samples)
  int frame_num, poc_lsb, idr_pic_id, nal_ref_idc;
  int slot;
  int serial;
  bool complete;                         // every macroblock was decoded
  bool output = true;                    // false: decoded before the first IDR picture of the stream, held back (see decoder.cpp `frozen`)
  bool idr = false;
  bool ltr_marking = false;              // the picture's marking made a long-term reference (MMCO 3 / 6): manage_dec_ref.cpp:376-378, :405-407
  int ltr_marked_frame_num = 0;
  unsigned long long timestamp;          // of the call that brought the picture's last slice ...
  unsigned long long timestamp_first = 0;// ... and its first one
  bool ref_lost = false, ref_lost_first = false;   // decoded after a lost reference picture (until the next IDR); the first such picture
  bool error = false;                    // concealed, or predicted from a picture that is not complete (!bIsComplete)
};

// Per-MB side record for the recompression model (filled by the parser; reference:
// DecodedMacroblock, codec/decoder/core/inc/decoded_macroblock.h:4-104).
struct ModelMbSide {
  int16_t mvd[16][2];        // as parsed, per 4x4 block in raster order (partition's first block and replicated)
  int8_t  pred_mode[16];     // signalled intra 4x4/8x8 modes (z order), before availability resolution
  int8_t  pred_mode_pred[16];// predicted mode for each block
  int16_t skip_run;          // mb_skip_run read before this MB (-1 if none)
  int8_t  qp_delta;
  uint8_t i16_mode_raw, chroma_mode_raw;
  uint8_t reserved;
};

// Where a Decoder writes the bulky hand-off records (MB records, coefficient pools, model records) of the picture it is parsing.
// Default: heap buffers of its own.  A host that stages records for the GPU (runtime.cpp) installs an arena so that the parser
// writes STRAIGHT into pinned staging memory: no second pass over ~2.6 MB per 1080p picture.
class RecordArena {
 public:
  enum Kind { MBS = 0, COEFS = 1, MRECS = 2 };
  virtual ~RecordArena() {}
  // Storage of at least `need` bytes of `kind` for the picture in progress; the `keep` bytes at `old` (this picture's records so
  // far; NULL when there are none) are preserved at the start of the returned block.  *cap = usable bytes.  NULL = no memory.
  virtual uint8_t* storage (int kind, size_t need, size_t* cap, const uint8_t* old, size_t keep) = 0;
};

class PictureSink {
 public:
  virtual ~PictureSink() {}
  // geometry changed (new SPS activated): (re)allocate picture storage with n_slots slots
  virtual int configure (int mb_w, int mb_h, int n_slots) = 0;
  // one complete coded picture in hand-off form; pointers valid until the call returns
  virtual int submit (const lh264_picture_desc& pic, const DecodedPictureInfo& info) = 0;
  // host-side error concealment: copy slot src -> dst (whole picture)
  virtual int copy_slot (int dst, int src) { (void)dst; (void)src; return 0; }
  // fill every sample of a slot with `value` (the grey stand-in for a reference picture that was never received)
  virtual int fill_slot (int slot, int value) { (void)slot; (void)value; return 0; }
  // the picture's recompression-model records (only when Decoder::enable_model(true)); pointers valid until return
  virtual int submit_model (const lh264_model_picture& pic, const DecodedPictureInfo& info) { (void)pic; (void)info; return 0; }
};

struct DecoderStats { uint64_t nals = 0, slices = 0, pictures = 0, mbs = 0, coef_blocks = 0, errors = 0;
                      uint64_t rewritten_slices = 0, rewrite_mismatches = 0; };   // enable_rewrite_check()

class Decoder {
 public:
  explicit Decoder (PictureSink* sink);
  ~Decoder();
  // Feeds bytes holding zero or more complete Annex-B NAL units.  Returns 0 or a negative error
  // for the LAST failing NAL (decoding continues with the next NAL).
  int decode (const uint8_t* data, size_t len, unsigned long long timestamp = 0);
  // Flushes a picture still pending (end of stream).
  int flush();
  const DecoderStats& stats() const { return stats_; }
  // self-check of the decompress direction: re-serialise every parsed slice_data() from its model records (cavlc_writer.h)
  // and compare with the bits that were parsed.  Implies enable_model().
  void enable_rewrite_check (bool on);
  // pictures completed by the last decode()/flush() call, in output order
  const std::vector<DecodedPictureInfo>& output() const { return out_; }
  // every picture finished by the last decode()/flush() call, handed out or not (what the reference calls a decoded access unit)
  const std::vector<DecodedPictureInfo>& finished() const { return fin_; }
  void clear_output() { out_.clear(); fin_.clear(); }
  void enable_model (bool on);           // produce lh264_model_picture hand-offs (PictureSink::submit_model)
  // Switches record storage to `arena` (NULL = the decoder's own heap buffers) for the following pictures AND for a picture in
  // progress, whose records so far move over.  The pointers PictureSink::submit() sees then point into the arena.
  bool set_arena (RecordArena* arena);
  bool picture_open() const;             // a picture has been started and not yet handed to the sink
  // non-temporal stores for the coefficient pools / the model records (worth it when they go to memory the parser never reads
  // back, e.g. pinned staging; not when the destination is cache-resident scratch).  Default: both on.
  void set_streaming_stores (bool coefs, bool model_recs);
  // The reference's FREEZE error-concealment modes (ERROR_CON_*_FREEZE_RES_CHANGE, what h264dec selects) hold the output back
  // from a lost IDR picture / a picture with missing macroblocks until the next complete IDR picture; the other modes hand every
  // picture out.  Default: freeze (the h264dec behaviour the YUV goldens pin).
  void set_freeze_output (bool on);
  // ERROR_CON_IDC as SDecodingParam::eEcActiveIdc (at_initialize) or DECODER_OPTION_ERROR_CON_IDC sets it; default: what h264dec
  // ends up with (Initialize with ERROR_CON_SLICE_COPY, then SetOption ERROR_CON_SLICE_MV_COPY_CROSS_IDR_FREEZE_RES_CHANGE)
  void set_ec_mode (int idc, bool at_initialize);
  // the picture being decoded predicts from a picture that is not complete (concealed / lost reference): bAllRefComplete false
  bool open_picture_error() const;
  // false: hand out every concealed picture (DecodeFrame2 semantics); true (default): as DecodeFrameNoDelay / h264dec does
  void set_nodelay_output (bool on);
  // the last slice NAL handled by the last decode() call (for the `.pip` default stream, pip_container.h); rbsp is valid
  // until the next decode() call
  struct LastSlice { bool decoded = false, cabac = false; const uint8_t* rbsp = nullptr; size_t header_bits = 0; bool idr = false; int idr_pic_id = 0, frame_num = 0; };
  const LastSlice& last_slice() const { return last_slice_; }
  bool sps_seen() const { return sps_seen_; }
  bool keep_model_side = false;          // fill ModelMbSide records
  const std::vector<ModelMbSide>& model_side() const;

 private:
  struct Impl;
  std::unique_ptr<Impl> d_;
  DecoderStats stats_;
  std::vector<DecodedPictureInfo> out_, fin_;
  LastSlice last_slice_;
  bool sps_seen_ = false;
};

} // namespace lh264host
