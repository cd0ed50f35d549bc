// openh264_api.cpp — libopenh264_lh264.so: the OpenH264 decoder API (include/lh264_codec_api.h)
// over the host decoder + GPU reconstruction.  Mirrors CWelsDecoder
// (codec/decoder/plus/src/welsDecoderExt.cpp:189-708): same entry points, argument meaning and
// return-code conventions; the macroblock reconstruction behind it runs on the B200 through the
// C ABI of include/lh264_b200.h (no CPU fallback: Initialize fails without a CUDA device).
#include <string.h>
#include <stdlib.h>
#include <deque>
#include <map>
#include <vector>
#include <new>
#include "../../include/lh264_codec_api.h"
#include "decoder.h"
#include "sinks.h"
#include "pip_container.h"
#include <string>
#include <stdio.h>

using namespace lh264host;

namespace {

class CLh264Decoder : public ISVCDecoder {
 public:
  CLh264Decoder() {}
  ~CLh264Decoder() override { Uninitialize(); }

  long Initialize (const SDecodingParam* p) override {
    if (!p) return cmInitParaError;
    if (p->bParseOnly) return cmUnsupportedData;
    Uninitialize();
    ec_idc_ = p->eEcActiveIdc;
    sink_ = new (std::nothrow) GpuSink (nullptr, device_);
    if (!sink_ || !sink_->ctx()) { delete sink_; sink_ = nullptr; return cmInitParaError; }   // no device: fail loudly
    dec_ = new (std::nothrow) Decoder (sink_);
    if (!dec_) return cmMallocMemeError;
    dec_->set_nodelay_output (false);                 // the two halves of DecodeFrameNoDelay are played out in decode()
    ec_applied_ = false;                              // the first mode comes with Initialize (InitErrorCon on a fresh context)
    apply_ec();
    memset (&stat_, 0, sizeof (stat_));
    return cmResultSuccess;
  }
  long Uninitialize() override {
    write_pip();
    delete dec_; dec_ = nullptr;
    delete sink_; sink_ = nullptr;
    if (out_) { lh264_host_free (out_); out_ = nullptr; out_cap_ = 0; }
    queue_.clear(); held_.clear(); saved_.clear();
    return cmResultSuccess;
  }

  DECODING_STATE DecodeFrame2 (const unsigned char* src, const int len, unsigned char** ppDst, SBufferInfo* info) override {
    return decode (src, len, ppDst, info, false);
  }
  // welsDecoderExt.cpp:408-430: DecodeFrame2(src) then DecodeFrame2(NULL, 0), return codes OR-ed
  DECODING_STATE DecodeFrameNoDelay (const unsigned char* src, const int len, unsigned char** ppDst, SBufferInfo* info) override {
    return decode (src, len, ppDst, info, true);
  }
  // welsDecoderExt.cpp:618-642
  DECODING_STATE DecodeFrame (const unsigned char* src, const int len, unsigned char** ppDst, int* stride, int& w, int& h) override {
    SBufferInfo bi;
    memset (&bi, 0, sizeof (bi));
    DECODING_STATE st = DecodeFrame2 (src, len, ppDst, &bi);
    if (bi.iBufferStatus == 1) {
      stride[0] = bi.UsrData.sSystemBuffer.iStride[0]; stride[1] = bi.UsrData.sSystemBuffer.iStride[1];
      w = bi.UsrData.sSystemBuffer.iWidth; h = bi.UsrData.sSystemBuffer.iHeight;
    }
    return st;
  }
  DECODING_STATE DecodeParser (const unsigned char*, const int, SParserBsInfo*) override { return dsInitialOptExpected; }
  DECODING_STATE DecodeFrameEx (const unsigned char*, const int, unsigned char*, int, int& iDstLen, int&, int&, int&) override { iDstLen = 0; return dsErrorFree; }

  long SetOption (DECODER_OPTION id, void* opt) override {
    if (id == DECODER_OPTION_LH264_PIP_OUTPUT) {       // extension: see include/lh264_codec_api.h
      if (!opt) return write_pip() ? cmResultSuccess : cmUnkonwReason;
      if (!dec_) return cmInitExpected;
      pip_path_ = (const char*)opt;
      pip_ = new (std::nothrow) PipBoilerplate();
      dec_->enable_model (true);
      // `.pipx`: the single-file form, written while the stream is being compressed (see flush_chunks)
      streaming_ = is_pip_archive_name (pip_path_.c_str());
      def_sent_ = 0; tag_sent_.assign (LH264_NUM_TAGS, 0); since_flush_ = 0;
      if (streaming_ && !arc_.open (pip_path_.c_str())) { delete pip_; pip_ = nullptr; return cmUnkonwReason; }
      return pip_ ? cmResultSuccess : cmMallocMemeError;
    }
    if (id == DECODER_OPTION_LH264_PIP_FLUSH_UNITS) {  // extension: units between two flushes of a `.pipx` (0: only at the end)
      if (!opt) return cmInitParaError;
      flush_units_ = * (int*)opt < 0 ? 0 : * (int*)opt;
      return cmResultSuccess;
    }
    if (!opt) return cmInitParaError;
    switch (id) {
    case DECODER_OPTION_DATAFORMAT: return * (int*)opt == videoFormatI420 ? cmResultSuccess : cmInitParaError;
    case DECODER_OPTION_END_OF_STREAM: eos_ = eos_opt_ = * (bool*)opt; return cmResultSuccess;
    case DECODER_OPTION_ERROR_CON_IDC: ec_idc_ = * (int*)opt; apply_ec(); return cmResultSuccess;
    case DECODER_OPTION_TRACE_LEVEL: trace_level_ = * (int*)opt; return cmResultSuccess;
    case DECODER_OPTION_TRACE_CALLBACK: case DECODER_OPTION_TRACE_CALLBACK_CONTEXT: return cmResultSuccess;
    default: return cmInitParaError;
    }
  }
  long GetOption (DECODER_OPTION id, void* opt) override {
    if (!opt) return cmInitParaError;
    switch (id) {
    case DECODER_OPTION_DATAFORMAT: * (int*)opt = videoFormatI420; return cmResultSuccess;
    case DECODER_OPTION_END_OF_STREAM: * (int*)opt = eos_; return cmResultSuccess;
    case DECODER_OPTION_VCL_NAL: * (int*)opt = last_vcl_; return cmResultSuccess;
    case DECODER_OPTION_TEMPORAL_ID: * (int*)opt = tid_; return cmResultSuccess;
    case DECODER_OPTION_FRAME_NUM: * (int*)opt = last_frame_num_; return cmResultSuccess;
    case DECODER_OPTION_IDR_PIC_ID: * (int*)opt = last_idr_pic_id_; return cmResultSuccess;
    case DECODER_OPTION_LTR_MARKING_FLAG: * (int*)opt = ltr_flag_ ? 1 : 0; return cmResultSuccess;      // (the reference stores an int, :370-373)
    case DECODER_OPTION_LTR_MARKED_FRAME_NUM: * (int*)opt = ltr_frame_num_; return cmResultSuccess;
    case DECODER_OPTION_ERROR_CON_IDC: * (int*)opt = ec_idc_; return cmResultSuccess;
    case DECODER_OPTION_GET_STATISTICS: memcpy (opt, &stat_, sizeof (stat_)); return cmResultSuccess;
    default: return cmInitParaError;
    }
  }

 private:
  // the two FREEZE modes hold output back behind a lost IDR / an incomplete picture (decoder_core.cpp:56-217 bFreezeOutput); the
  // others hand every picture out
  void apply_ec() {
    if (dec_) { dec_->set_ec_mode (ec_idc_, !ec_applied_); ec_applied_ = true; }
  }
  // One DecodeFrame2 call as an API client sees it (welsDecoderExt.cpp:431-571).  The reference decodes an access unit when the
  // first NAL of the NEXT one arrives, or at once on the flush call DecodeFrame2 (NULL, 0) — which DecodeFrameNoDelay issues after
  // every unit (:408-430).  The host decoder finishes a complete picture at its last slice, so DecodeFrame2 holds such a picture
  // back until the next call: output timing, time stamps and the per-call feedback options then follow the reference.
  DECODING_STATE decode_one (const unsigned char* src, int len, unsigned char** ppDst, SBufferInfo* info, bool second_half = false) {
    if (!dec_ || !sink_) return dsInitialOptExpected;
    if (!ppDst || !info) return dsInvalidArgument;
    const unsigned long long ts = info->uiInBsTimeStamp;       // welsDecoderExt.cpp:463-465
    memset (info, 0, sizeof (*info));
    info->uiInBsTimeStamp = ts;
    ppDst[0] = ppDst[1] = ppDst[2] = nullptr;
    int state = dsErrorFree;
    // feedback of this call (welsDecoderExt.cpp:461-473)
    last_vcl_ = FEEDBACK_UNKNOWN_NAL; tid_ = -1; last_frame_num_ = -1; ltr_flag_ = false; ltr_frame_num_ = 0;
    std::vector<DecodedPictureInfo> decoded_now;               // access units the reference would decode during this call
    dec_->clear_output();
    if (src && len > 0) {
      eos_ = false;
      for (const DecodedPictureInfo& o : held_) decoded_now.push_back (o);     // the next access unit has arrived
      held_.clear();
      const int rc = dec_->decode (src, (size_t)len, ts);
      if (dec_->last_slice().decoded && dec_->last_slice().idr) last_idr_pic_id_ = dec_->last_slice().idr_pic_id;
      if (rc == -2) state |= dsNoParamSets;
      else if (rc < 0) state |= dsBitstreamError;
      // finished by its own last slice: the reference has not decoded it yet; finished by the arrival of this unit: it has
      for (const DecodedPictureInfo& o : dec_->finished()) {
        if (!o.complete) { decoded_now.push_back (o); continue; }
        held_.push_back (o);
        // the picture is handed out at a LATER call, after more data has been decoded: a picture that is no reference may have lost
        // its slot to the next one by then, so its samples are taken now (DecodeFrame2 only; the flush half of DecodeFrameNoDelay
        // releases it before anything else is decoded)
        if (!nodelay_ && o.output) {
          const int cw = (o.width + 1) >> 1, ch = (o.height + 1) >> 1;
          std::vector<unsigned char>& px = saved_[o.serial];
          px.resize ((size_t)o.width * o.height + 2 * (size_t)cw * ch);
          if (sink_->read (o, px.data(), o.width, px.data() + (size_t)o.width * o.height, px.data() + (size_t)o.width * o.height + (size_t)cw * ch, cw) != LH264_OK) saved_.erase (o.serial);
        }
      }
    } else {
      eos_ = true;                                               // :452-455: the flush call
      // a picture still missing slices is only given up at the END of the stream (the client said so with DECODER_OPTION_END_OF_STREAM),
      // not by the flush half of DecodeFrameNoDelay, which comes after every slice of a multi-slice picture
      if (eos_opt_ && !second_half) dec_->flush();
      for (const DecodedPictureInfo& o : held_) decoded_now.push_back (o);
      held_.clear();
      for (const DecodedPictureInfo& o : dec_->finished()) decoded_now.push_back (o);
    }
    for (const DecodedPictureInfo& o : decoded_now) {
      last_vcl_ = FEEDBACK_VCL_NAL; tid_ = 0; last_frame_num_ = o.frame_num;
      if (o.ltr_marking) { ltr_flag_ = true; ltr_frame_num_ = o.ltr_marked_frame_num; }
      if (o.output) queue_.push_back (o);
      if (!o.complete || !o.output || o.ref_lost || o.error) state |= dsDataErrorConcealed;   // (incomplete, held back behind a lost IDR picture, or predicted from a lost one)
      if (o.ref_lost_first) state |= dsRefLost;
    }
    if (!queue_.empty()) {
      const DecodedPictureInfo o = queue_.front();
      queue_.pop_front();
      const int cw = (o.width + 1) >> 1, ch = (o.height + 1) >> 1;
      const size_t need = (size_t)o.width * o.height + 2 * (size_t)cw * ch;
      if (need > out_cap_) {
        if (out_) lh264_host_free (out_);
        out_ = (unsigned char*)lh264_host_alloc (need);
        out_cap_ = out_ ? need : 0;
        if (!out_) return (DECODING_STATE) (state | dsOutOfMemory);
      }
      unsigned char* y = out_, *u = y + (size_t)o.width * o.height, *v = u + (size_t)cw * ch;
      auto sv = saved_.find (o.serial);
      if (sv != saved_.end()) { memcpy (y, sv->second.data(), need); saved_.erase (sv); }
      else if (sink_->read (o, y, o.width, u, v, cw) != LH264_OK) return (DECODING_STATE) (state | dsOutOfMemory);
      ppDst[0] = y; ppDst[1] = u; ppDst[2] = v;
      info->iBufferStatus = 1;
      // the time stamp that came with the picture's own data: DecodeFrame2 decodes an access unit as a whole (its first unit's
      // stamp), the flush half of DecodeFrameNoDelay slice by slice (the last one's)
      info->uiOutYuvTimeStamp = nodelay_ ? o.timestamp : o.timestamp_first;
      info->UsrData.sSystemBuffer.iWidth = o.width; info->UsrData.sSystemBuffer.iHeight = o.height;
      info->UsrData.sSystemBuffer.iFormat = videoFormatI420;
      info->UsrData.sSystemBuffer.iStride[0] = o.width; info->UsrData.sSystemBuffer.iStride[1] = cw;
      stat_.uiDecodedFrameCount++;
      if (stat_.uiWidth != (unsigned)o.width || stat_.uiHeight != (unsigned)o.height) { stat_.uiResolutionChangeTimes++; stat_.uiWidth = o.width; stat_.uiHeight = o.height; }
    }
    return (DECODING_STATE)state;
  }

  DECODING_STATE decode (const unsigned char* src, int len, unsigned char** ppDst, SBufferInfo* info, bool no_delay) {
    nodelay_ = no_delay;
    if (!no_delay) return decode_one (src, len, ppDst, info);
    if (!dec_ || !sink_) return dsInitialOptExpected;
    if (!ppDst || !info) return dsInvalidArgument;
    if (pip_ && src && len > 0) {
      if (len < 4) { pip_->raw (src, (size_t)len); return dsErrorFree; }     // h264dec.cpp:268-271
      pip_->unit_begin (src, (size_t)len, dec_->sps_seen());
    }
    // DecodeFrame2 (src) then DecodeFrame2 (NULL, 0), return codes OR-ed; the caller sees the second call's output (:408-430)
    int st = (int)decode_one (src, len, ppDst, info);
    if (pip_ && src && len > 0) {
      UnitSliceInfo u;
      u.slice_decoded = dec_->last_slice().decoded; u.cabac = dec_->last_slice().cabac;
      u.rbsp = dec_->last_slice().rbsp; u.header_bits = dec_->last_slice().header_bits;
      pip_->unit_end (u);
    }
    unsigned char* first_planes[3] = {ppDst[0], ppDst[1], ppDst[2]};
    const SBufferInfo first = *info;
    const bool slice_here = src && len > 0 && dec_->last_slice().decoded;
    const int slice_frame_num = dec_->last_slice().frame_num;
    st |= (int)decode_one (nullptr, 0, ppDst, info, true);
    if (pip_ && streaming_ && flush_units_ > 0 && ++since_flush_ >= flush_units_) { since_flush_ = 0; flush_chunks (false); }
    // the flush half decodes what the unit brought at once: a slice unit reports its access unit even while the picture waits for
    // more slices
    if (slice_here && last_vcl_ != FEEDBACK_VCL_NAL) { last_vcl_ = FEEDBACK_VCL_NAL; tid_ = 0; last_frame_num_ = slice_frame_num; }
    // A picture released by the FIRST half (a concealed one, given up when the next access unit arrived) is lost to the caller:
    // the second DecodeFrame2 clears the output structure again (welsDecoderExt.cpp:408-430; the lines that would have kept it are
    // commented out in the reference) - `first` is not looked at.
    (void)first; (void)first_planes;
    // a slice decoded at once (flush half) from a picture that is not complete reports the concealment state with every call
    if (slice_here && dec_->open_picture_error()) st |= dsDataErrorConcealed;
    return (DECODING_STATE)st;
  }

  // Streaming flush (the reference writes everything at exit, FlushOnClose h264dec.cpp:105-121): appends to the
  // `.pipx` what has become FINAL since the last call.  The default stream only grows.  A tag stream is the output of a VP8 bool
  // coder, whose pending carry can still add one to the number formed by the bytes written so far - that changes the last byte
  // that is not 0xff and the 0xff bytes behind it, nothing in front of it - so everything in front of the last non-0xff byte
  // goes out and the rest waits for the next call (`final`: the coders have been ended, everything goes).
  bool flush_chunks (bool final) {
    bool ok = true;
    const std::vector<uint8_t>& d = final ? pip_->finish() : pip_->so_far();
    if (d.size() > def_sent_) { ok = arc_.chunk (0xffffffffu, d.data() + def_sent_, d.size() - def_sent_) && ok; def_sent_ = d.size(); }
    if (!final && sink_->sync() != LH264_OK) return false;        // the deferred coder has coded every picture submitted so far
    std::vector<uint8_t> b;
    for (int tag = 1; tag < LH264_NUM_TAGS && ok; tag++) {
      if (sink_->model_tag_bytes_from (tag, tag_sent_[tag], b) != LH264_OK) return false;
      size_t n = b.size();
      if (!final) {
        while (n > 0 && b[n - 1] == 0xff) n--;          // trailing 0xff bytes and ...
        if (n > 0) n--;                                 // ... the byte in front of them may still change
      }
      if (n > 0) { ok = arc_.chunk ((uint32_t)tag, b.data(), n) && ok; tag_sent_[tag] += n; }
    }
    return ok;
  }

  // writes the `.pip` file set: default stream + one file per tag that has a stream (MultiFileWriter, h264dec.cpp:79-104)
  bool write_pip() {
    if (!pip_ || !sink_ || !dec_) return true;
    dec_->flush();
    bool ok = sink_->model_finish() == LH264_OK;
    if (streaming_) {
      ok = flush_chunks (true) && ok;
      ok = arc_.close() && ok;
      delete pip_; pip_ = nullptr;
      return ok;
    }
    const std::vector<uint8_t>& d = pip_->finish();
    FILE* f = fopen (pip_path_.c_str(), "wb");
    if (f) { fwrite (d.data(), 1, d.size(), f); fclose (f); } else ok = false;
    std::vector<uint8_t> b;
    for (int tag = 1; tag < LH264_NUM_TAGS && ok; tag++) {
      if (sink_->model_tag_bytes (tag, b) != LH264_OK) { ok = false; break; }
      if (b.empty()) continue;
      const std::string name = pip_path_ + "." + std::to_string (tag);
      FILE* ft = fopen (name.c_str(), "wb");
      if (ft) { fwrite (b.data(), 1, b.size(), ft); fclose (ft); } else ok = false;
    }
    delete pip_; pip_ = nullptr;
    return ok;
  }

  Decoder* dec_ = nullptr;
  GpuSink* sink_ = nullptr;
  PipBoilerplate* pip_ = nullptr;
  std::string pip_path_;
  bool streaming_ = false;                        // `.pipx` output: chunks go out every flush_units_ units
  PipStreamWriter arc_;
  size_t def_sent_ = 0;
  std::vector<size_t> tag_sent_;
  int flush_units_ = 64, since_flush_ = 0;
  std::deque<DecodedPictureInfo> queue_;
  std::vector<DecodedPictureInfo> held_;          // complete pictures the reference would only decode at the next call
  std::map<int, std::vector<unsigned char> > saved_;   // ... and their samples (by DecodedPictureInfo::serial)
  int tid_ = -1, ltr_frame_num_ = 0;
  bool ltr_flag_ = false, nodelay_ = false;
  unsigned char* out_ = nullptr;
  size_t out_cap_ = 0;
  bool eos_ = false, eos_opt_ = false;
  int ec_idc_ = ERROR_CON_SLICE_COPY, trace_level_ = 0, device_ = 0;
  bool ec_applied_ = false;
  int last_vcl_ = 0, last_frame_num_ = -1, last_idr_pic_id_ = 0;
  SDecoderStatistics stat_;
};

} // namespace

extern "C" {

long WelsCreateDecoder (ISVCDecoder** pp) {
  if (!pp) return 1;                                  // ERR_INVALID_PARAMETERS
  *pp = new (std::nothrow) CLh264Decoder();
  return *pp ? 0 : 2;
}
void WelsDestroyDecoder (ISVCDecoder* p) { delete static_cast<CLh264Decoder*> (p); }
int WelsGetDecoderCapability (SDecoderCapability* c) {   // fixed values, welsDecoderExt.cpp:666-679
  if (!c) return 1;
  memset (c, 0, sizeof (*c));
  c->iProfileIdc = 66; c->iProfileIop = 0xE0; c->iLevelIdc = 32; c->iMaxMbps = 216000; c->iMaxFs = 5120;
  c->iMaxCpb = 20000; c->iMaxDpb = 20480; c->iMaxBr = 20000; c->bRedPicCap = 0;
  return 0;
}
OpenH264Version WelsGetCodecVersion (void) { OpenH264Version v = {1, 4, 0, 0}; return v; }
void WelsGetCodecVersionEx (OpenH264Version* v) { if (v) *v = WelsGetCodecVersion(); }

} // extern "C"
