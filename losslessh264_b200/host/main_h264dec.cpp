// main_h264dec.cpp — h264dec-compatible console decoder over libopenh264_lh264.so.
// Mirrors the reference CLI contract (codec/console/dec/src/h264dec.cpp:123-402, :407-564):
//   h264dec <in.264> <out.yuv>           decode, write cropped I420 frames
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

static void write_frame (FILE* f, unsigned char** d, const SBufferInfo& bi) {
  const int w = bi.UsrData.sSystemBuffer.iWidth, h = bi.UsrData.sSystemBuffer.iHeight;
  const int sy = bi.UsrData.sSystemBuffer.iStride[0], sc = bi.UsrData.sSystemBuffer.iStride[1];
  for (int y = 0; y < h; y++) fwrite (d[0] + (size_t)y * sy, 1, w, f);
  for (int p = 1; p < 3; p++) for (int y = 0; y < h / 2; y++) fwrite (d[p] + (size_t)y * sc, 1, w / 2, f);
}

int main (int argc, char** argv) {
  if (argc < 3) { fprintf (stderr, "usage: %s <in.264> <out.yuv>\n", argv[0]); return 1; }
  if (strstr (argv[2], ".pip") || strstr (argv[1], ".pip")) { fprintf (stderr, "the .pip recompression modes are not built into this binary yet\n"); return 2; }
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
  FILE* fo = fopen (argv[2], "wb");
  if (!fo) { perror (argv[2]); return 1; }

  ISVCDecoder* dec = nullptr;
  if (WelsCreateDecoder (&dec) || !dec) { fprintf (stderr, "WelsCreateDecoder failed\n"); return 1; }
  SDecodingParam param;
  memset (&param, 0, sizeof (param));
  param.eOutputColorFormat = videoFormatI420; param.uiTargetDqLayer = 255; param.eEcActiveIdc = ERROR_CON_SLICE_COPY;
  param.sVideoProperty.eVideoBsType = VIDEO_BITSTREAM_DEFAULT;
  if (dec->Initialize (&param)) { fprintf (stderr, "Initialize failed (no CUDA device? this build has no CPU fallback)\n"); return 1; }

  int frames = 0, width = 0, height = 0;
  double secs = 0;
  // start-code positions (a 4-byte start code's leading zero is kept with the NAL it introduces)
  std::vector<size_t> starts;
  for (size_t i = 0; i + 3 <= size + 4; i++)
    if (buf[i] == 0 && buf[i + 1] == 0 && buf[i + 2] == 1) { starts.push_back (i > 0 && buf[i - 1] == 0 && (starts.empty() || starts.back() + 3 < i) ? i - 1 : i); i += 2; }
  for (size_t k = 0; k + 1 < starts.size(); k++) {
    const size_t pos = starts[k], len = starts[k + 1] - starts[k];
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
