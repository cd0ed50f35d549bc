// main_h264dec.cpp — h264dec-compatible console decoder over libopenh264_lh264.so.
// Mirrors the reference CLI contract (codec/console/dec/src/h264dec.cpp:123-402, :407-564):
//   h264dec <in.264> <out.yuv>           decode, write cropped I420 frames
//   h264dec <in.264> <out.pip> [yuv]     compress: decode + recompression model; writes out.pip and out.pip.<tag>
//   h264dec <in.pip> <out.264>           decompress: in.pip + in.pip.<tag> -> the original Annex-B file, byte for byte
//   ... and `.pipx` in place of `.pip`: the single-file container (pip_container.h), an addition of this build
// one NAL unit per DecodeFrameNoDelay call (3- or 4-byte start codes), the same getters queried
// after every call, the same "iWidth/height/Frames/decode time/FPS" block on stderr.
// The .pip compress/decompress modes are selected by file-name suffix exactly like the reference
// (:167-177) and are served by the recompression layer when it is built in.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <vector>
#include "../../include/lh264_codec_api.h"
#include "pip_container.h"
#include "pip_expand.h"
#include <string>

static void write_frame (FILE* f, unsigned char** d, const SBufferInfo& bi) {
  const int w = bi.UsrData.sSystemBuffer.iWidth, h = bi.UsrData.sSystemBuffer.iHeight;
  const int sy = bi.UsrData.sSystemBuffer.iStride[0], sc = bi.UsrData.sSystemBuffer.iStride[1];
  for (int y = 0; y < h; y++) fwrite (d[0] + (size_t)y * sy, 1, w, f);
  for (int p = 1; p < 3; p++) for (int y = 0; y < h / 2; y++) fwrite (d[p] + (size_t)y * sc, 1, w / 2, f);
}

static bool slurp (const std::string& path, std::vector<uint8_t>& out) {
  FILE* f = fopen (path.c_str(), "rb");
  if (!f) return false;
  uint8_t tmp[65536];
  size_t n;
  out.clear();
  while ((n = fread (tmp, 1, sizeof (tmp), f)) > 0) out.insert (out.end(), tmp, tmp + n);
  fclose (f);
  return true;
}

// Round-trip mode of the reference CLI (isRecoding, h264dec.cpp:170-173): the default stream gives the NAL skeleton and the
// slice headers, the side streams go through the model on the GPU (lh264_model_decode), the host frames the result.
static int decompress (const char* pip_path, const char* out_path) {
  lh264host::PipFileSet fs;
  if (!(lh264host::is_pip_archive_name (pip_path) ? lh264host::pip_archive_read (pip_path, fs) : lh264host::pip_set_read (pip_path, fs))) { fprintf (stderr, "cannot read %s\n", pip_path); return 1; }
  std::vector<uint8_t>& def = fs.def;
  std::vector<std::vector<uint8_t> >& tags = fs.tags;
  lh264_decode_job job;
  memset (&job, 0, sizeof (job));
  size_t total = def.size();
  for (int t = 0; t < LH264_NUM_TAGS; t++)
    if (fs.present[t]) {
      tags[t].push_back (0);
      job.tag_data[t] = tags[t].data(); job.tag_len[t] = (uint32_t)tags[t].size() - 1; total += tags[t].size();
    }
  lh264host::ExpandPlan plan;
  if (!lh264host::expand_plan (def.data(), def.size(), plan)) { fprintf (stderr, "cannot expand %s: %s\n", pip_path, plan.error.c_str()); return 3; }
  lh264_ctx* ctx = nullptr;
  if (lh264_ctx_create (&ctx, 0, nullptr) != LH264_OK) { fprintf (stderr, "%s (this build has no CPU fallback)\n", lh264_last_error()); return 1; }
  std::vector<lh264_decode_slice_out> outs (plan.slices.size() + 1);
  std::vector<uint8_t> bits;
  const auto t0 = std::chrono::steady_clock::now();
  int rc = LH264_OK;
  for (size_t cap = total * 4 + (1u << 20);; cap *= 2) {
    bits.assign (cap, 0);
    job.max_mbs = plan.max_mbs > 0 ? plan.max_mbs : 1; job.with_8x8 = plan.with_8x8;
    job.slices = plan.slices.data(); job.n_slices = (int)plan.slices.size();
    job.out = bits.data(); job.out_capacity = bits.size(); job.slice_out = outs.data();
    rc = lh264_model_decode (ctx, &job, 1);
    if (job.error != 22) break;                        // 22: output buffer too small
  }
  const double secs = std::chrono::duration<double> (std::chrono::steady_clock::now() - t0).count();
  if (rc != LH264_OK) { fprintf (stderr, "model decode failed: %s\n", lh264_last_error()); lh264_ctx_destroy (ctx); return 4; }
  std::vector<uint8_t> out;
  const float kms = lh264_kernel_ms (ctx, 6);
  const bool ok = lh264host::expand_assemble (plan, def.data(), bits.data(), outs.data(), out);
  lh264_ctx_destroy (ctx);
  if (!ok) { fprintf (stderr, "assembly failed\n"); return 5; }
  FILE* fo = fopen (out_path, "wb");
  if (!fo) { perror (out_path); return 1; }
  fwrite (out.data(), 1, out.size(), fo);
  fclose (fo);
  fprintf (stderr, "-------------------------------------------------------\n");
  fprintf (stderr, "slices:\t\t%zu\noutput bytes:\t%zu\nmodel decode time:\t%f sec (k_model_decode %.3f ms)\n", plan.slices.size(), out.size(), secs, (double)kms);
  fprintf (stderr, "-------------------------------------------------------\n");
  return 0;
}

int main (int argc, char** argv) {
  if (argc < 3) { fprintf (stderr, "usage: %s <in.264> <out.yuv> | <in.264> <out.pip> [out.yuv] | <in.pip> <out.264>\n", argv[0]); return 1; }
  const bool compress = strstr (argv[2], ".pip") != nullptr;             // h264dec.cpp:167-169
  if (!compress && strstr (argv[1], ".pip")) return decompress (argv[1], argv[2]);                        // h264dec.cpp:170-173
  const char* yuv_path = compress ? (argc > 3 ? argv[3] : "/dev/null") : argv[2];
  // `.pipx`: the decoder object writes the single-file container itself, chunk by chunk while it compresses
  const std::string set_path = std::string (argv[2]);
  FILE* fi = fopen (argv[1], "rb");
  if (!fi) { perror (argv[1]); return 1; }
  std::vector<unsigned char> buf;
  unsigned char tmp[65536];
  size_t n;
  while ((n = fread (tmp, 1, sizeof (tmp), fi)) > 0) buf.insert (buf.end(), tmp, tmp + n);
  fclose (fi);
  const size_t size = buf.size();
  static const unsigned char kStart[4] = {0, 0, 0, 1};
  buf.insert (buf.end(), kStart, kStart + 4);          // sentinel, h264dec.cpp:227-238
  FILE* fo = fopen (yuv_path, "wb");
  if (!fo) { perror (yuv_path); return 1; }

  ISVCDecoder* dec = nullptr;
  if (WelsCreateDecoder (&dec) || !dec) { fprintf (stderr, "WelsCreateDecoder failed\n"); return 1; }
  SDecodingParam param;
  memset (&param, 0, sizeof (param));
  param.eOutputColorFormat = videoFormatI420; param.uiTargetDqLayer = 255; param.eEcActiveIdc = ERROR_CON_SLICE_COPY;
  param.sVideoProperty.eVideoBsType = VIDEO_BITSTREAM_DEFAULT;
  if (dec->Initialize (&param)) { fprintf (stderr, "Initialize failed (no CUDA device? this build has no CPU fallback)\n"); return 1; }
  { int ec = ERROR_CON_SLICE_MV_COPY_CROSS_IDR_FREEZE_RES_CHANGE; dec->SetOption (DECODER_OPTION_ERROR_CON_IDC, &ec); }   // h264dec.cpp:156-157
  if (compress && dec->SetOption (DECODER_OPTION_LH264_PIP_OUTPUT, (void*)set_path.c_str())) { fprintf (stderr, "cannot switch the recompressor on\n"); return 1; }
  if (const char* e = getenv ("LH264_PIP_FLUSH_UNITS")) { int n = atoi (e); dec->SetOption (DECODER_OPTION_LH264_PIP_FLUSH_UNITS, &n); }

  int frames = 0, width = 0, height = 0;
  double secs = 0;
  // the Annex-B unit loop of the reference CLI (h264dec.cpp:246-364): a unit runs up to the next 3- or 4-byte start code
  for (size_t pos = 0; pos < size;) {
    const size_t len = lh264host::PipBoilerplate::next_unit (buf.data(), pos, size);
    struct Advance { size_t& p; size_t n; ~Advance() { p += n; } } advance{pos, len};
    unsigned char* dst[3] = {nullptr, nullptr, nullptr};
    SBufferInfo bi;
    memset (&bi, 0, sizeof (bi));
    const auto t0 = std::chrono::steady_clock::now();
    dec->DecodeFrameNoDelay (buf.data() + pos, (int)len, dst, &bi);
    secs += std::chrono::duration<double> (std::chrono::steady_clock::now() - t0).count();
    int v;
    dec->GetOption (DECODER_OPTION_FRAME_NUM, &v); dec->GetOption (DECODER_OPTION_IDR_PIC_ID, &v); dec->GetOption (DECODER_OPTION_VCL_NAL, &v);
    if (bi.iBufferStatus == 1) { write_frame (fo, dst, bi); frames++; width = bi.UsrData.sSystemBuffer.iWidth; height = bi.UsrData.sSystemBuffer.iHeight; }
  }
  // drain: end of stream
  bool eos = true;
  dec->SetOption (DECODER_OPTION_END_OF_STREAM, &eos);
  for (;;) {
    unsigned char* dst[3] = {nullptr, nullptr, nullptr};
    SBufferInfo bi;
    memset (&bi, 0, sizeof (bi));
    dec->DecodeFrame2 (nullptr, 0, dst, &bi);
    if (bi.iBufferStatus != 1) break;
    write_frame (fo, dst, bi); frames++;
  }
  fclose (fo);
  fprintf (stderr, "-------------------------------------------------------\n");
  fprintf (stderr, "iWidth:\t\t%d\nheight:\t\t%d\nFrames:\t\t%d\ndecode time:\t%f sec\nFPS:\t\t%f fps\n", width, height, frames, secs, frames / (secs > 0 ? secs : 1));
  fprintf (stderr, "-------------------------------------------------------\n");
  dec->Uninitialize();
  WelsDestroyDecoder (dec);
  return 0;
}
