// api_probe.cpp — TEST INFRASTRUCTURE.  One source, two builds, to pin the OpenH264 decoder API behaviour at the boundary:
//   -DPROBE_REFERENCE : against the UNMODIFIED reference (its headers, its objects: oracle/Makefile target _ref/api_probe_ref);
//                       its output for a few streams is committed under tests/golden/api_probe/ (oracle/make_api_probe_golden.py)
//   default           : against include/lh264_codec_api.h + libopenh264_lh264.so (built by tests/test_api_boundary.py on the GPU box)
// It drives a decoder the way the reference's h264dec does (h264dec.cpp:246-364: one call per Annex-B unit, the GetOption reads
// of :275-290 in front of every call) and prints, per call, what an API client can observe: the return state, whether a picture
// came out and its size, and the option values.
//   api_probe nodelay|frame2 in.264 [error concealment idc, set with SetOption after Initialize as h264dec.cpp:156-157 does]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#ifdef PROBE_REFERENCE
#include "codec_def.h"
#include "codec_app_def.h"
#include "codec_api.h"
#include "compression_stream.h"
#else
#include "lh264_codec_api.h"
#endif

int main (int argc, char** argv) {
  if (argc < 3) { fprintf (stderr, "usage: %s nodelay|frame2 in.264\n", argv[0]); return 2; }
  const bool nodelay = !strcmp (argv[1], "nodelay");
  FILE* f = fopen (argv[2], "rb");
  if (!f) { perror (argv[2]); return 1; }
  std::vector<unsigned char> buf;
  unsigned char tmp[65536];
  size_t n;
  while ((n = fread (tmp, 1, sizeof (tmp), f)) > 0) buf.insert (buf.end(), tmp, tmp + n);
  fclose (f);
  const size_t size = buf.size();
  static const unsigned char kStart[4] = {0, 0, 0, 1};
  buf.insert (buf.end(), kStart, kStart + 4);
#ifdef PROBE_REFERENCE
  oMovie().isRecoding = false;                         // h264dec.cpp:169, :218-219
  iMovie().setOriginalFileSize ((int)size);
#endif
  ISVCDecoder* dec = NULL;
  if (WelsCreateDecoder (&dec) != 0 || !dec) { fprintf (stderr, "WelsCreateDecoder failed\n"); return 3; }
  SDecodingParam prm;
  memset (&prm, 0, sizeof (prm));
  prm.eOutputColorFormat = videoFormatI420;
  prm.uiTargetDqLayer = (unsigned char) - 1;
  prm.eEcActiveIdc = ERROR_CON_SLICE_COPY;
  prm.sVideoProperty.eVideoBsType = VIDEO_BITSTREAM_DEFAULT;
  if (dec->Initialize (&prm) != 0) { fprintf (stderr, "Initialize failed\n"); return 4; }
  if (argc > 3) { int ec = atoi (argv[3]); dec->SetOption (DECODER_OPTION_ERROR_CON_IDC, &ec); }
  int call = 0;
  auto options = [&] (const char* when) {
    int fmt = 0, eos = 0, idr = 0, fn = 0, ltr = 0, ltr_fn = 0, vcl = 0, tid = 0, ec = 0;   // (bool-typed options write one byte)
    dec->GetOption (DECODER_OPTION_DATAFORMAT, &fmt);
    dec->GetOption (DECODER_OPTION_END_OF_STREAM, &eos);
    dec->GetOption (DECODER_OPTION_IDR_PIC_ID, &idr);
    dec->GetOption (DECODER_OPTION_FRAME_NUM, &fn);
    dec->GetOption (DECODER_OPTION_LTR_MARKING_FLAG, &ltr);
    dec->GetOption (DECODER_OPTION_LTR_MARKED_FRAME_NUM, &ltr_fn);
    dec->GetOption (DECODER_OPTION_VCL_NAL, &vcl);
    dec->GetOption (DECODER_OPTION_TEMPORAL_ID, &tid);
    dec->GetOption (DECODER_OPTION_ERROR_CON_IDC, &ec);
    printf ("%s %d: fmt %d eos %d idr_pic_id %d frame_num %d ltr %d ltr_frame_num %d vcl %d tid %d ec %d\n", when, call, fmt, eos & 1, idr, fn, ltr & 1, ltr_fn, vcl, tid, ec);
  };
  auto one = [&] (const unsigned char* src, int len) {
    unsigned char* planes[3] = {NULL, NULL, NULL};
    SBufferInfo bi;
    memset (&bi, 0, sizeof (bi));
    bi.uiInBsTimeStamp = (unsigned long long) (call + 1);
    const int st = nodelay && src ? (int)dec->DecodeFrameNoDelay (src, len, planes, &bi) : (int)dec->DecodeFrame2 (src, len, planes, &bi);
    printf ("call %d: len %d state %d status %d", call, len, st, bi.iBufferStatus);
    // (strides are the library's business - the reference hands out its padded planes - and are not printed)
    if (bi.iBufferStatus == 1) {
      printf (" %dx%d format %d ts %llu", bi.UsrData.sSystemBuffer.iWidth, bi.UsrData.sSystemBuffer.iHeight, bi.UsrData.sSystemBuffer.iFormat, bi.uiOutYuvTimeStamp);
      // FNV-1a over the cropped planes: the SAMPLES an API client receives (concealed pictures included)
      unsigned long long h = 1469598103934665603ull;
      const int w = bi.UsrData.sSystemBuffer.iWidth, hh = bi.UsrData.sSystemBuffer.iHeight;
      for (int p = 0; p < 3; p++) {
        const int pw = p ? w / 2 : w, ph = p ? hh / 2 : hh, st = bi.UsrData.sSystemBuffer.iStride[p ? 1 : 0];
        for (int r = 0; r < ph; r++) for (int x = 0; x < pw; x++) { h ^= planes[p][(size_t)r * st + x]; h *= 1099511628211ull; }
      }
      printf (" samples %016llx", h);
    }
    printf ("\n");
    options ("after");
    call++;
  };
  // the unit loop of h264dec.cpp:246-267: a unit ends where the next 00 00 00 01 / 00 00 01 begins
  for (size_t pos = 0; pos < size;) {
    size_t i = 0;
    const unsigned char* p = buf.data() + pos;
    for (i = 0; i < size - pos; i++)
      if ((p[i] == 0 && p[i + 1] == 0 && p[i + 2] == 0 && p[i + 3] == 1 && i > 0) || (p[i] == 0 && p[i + 1] == 0 && p[i + 2] == 1 && i > 0)) break;
    const size_t len = i;
    if (len < 4) { pos += len; continue; }
    one (p, (int)len);
    pos += len;
  }
  int eos = 1;
  dec->SetOption (DECODER_OPTION_END_OF_STREAM, &eos);
  one (NULL, 0);
  dec->Uninitialize();
  WelsDestroyDecoder (dec);
  return 0;
}
