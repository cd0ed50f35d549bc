// api_driver.cpp — the reference's API-level decoder test procedure against libopenh264_lh264.so:
//   test/api/BaseDecoderTest.cpp:8-43   ReadFrame: the file is cut at every 4-byte start code 00 00 00 01
//   :49-65                              WelsCreateDecoder + Initialize (I420, uiTargetDqLayer = UCHAR_MAX, ERROR_CON_SLICE_COPY)
//   :74-110                             DecodeFrame2 per unit, every output frame to the callback
//   :111-133                            DECODER_OPTION_END_OF_STREAM, then DecodeFrame2 (NULL, 0) for the pending last frame
//   test/api/decoder_test.cpp:6-12,63-70  SHA-1 over the w x h rows of each plane (stride skipped), planes Y, U, V
// usage: api_driver vtable            -> C / C++ vtable order + layout check (no device needed), exit code 0 = ok
//        api_driver decode in.264     -> prints "frames N sha1 HEX state_or MASK"
//        api_driver cinit             -> a decoder created, initialised and destroyed from C through the function table
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <string>
#include <vector>
#include "../../include/lh264_codec_api.h"
#include "../../losslessh264_b200/host/sha1.h"

extern "C" {
// c_vtable_check.c sees ISVCDecoder as `const ISVCDecoderVtbl*`; the object pointer is the same
int lh264_check_decoder_vtable_c (void* p);
size_t lh264_c_bool_size (void);
size_t lh264_c_bool_offset (void);
size_t lh264_c_buffer_info_size (void);
size_t lh264_c_decoding_param_size (void);
long lh264_c_create_init_destroy (void);
}

namespace {
struct FakeDecoder : public ISVCDecoder {      // every entry returns its ordinal (test/api/cpp_interface_test.cpp:64-110)
  long Initialize (const SDecodingParam*) override { return 1; }
  long Uninitialize() override { return 2; }
  DECODING_STATE DecodeFrame (const unsigned char*, const int, unsigned char**, int*, int&, int&) override { return (DECODING_STATE)3; }
  DECODING_STATE DecodeFrameNoDelay (const unsigned char*, const int, unsigned char**, SBufferInfo*) override { return (DECODING_STATE)4; }
  DECODING_STATE DecodeFrame2 (const unsigned char*, const int, unsigned char**, SBufferInfo*) override { return (DECODING_STATE)5; }
  DECODING_STATE DecodeParser (const unsigned char*, const int, SParserBsInfo*) override { return (DECODING_STATE)6; }
  DECODING_STATE DecodeFrameEx (const unsigned char*, const int, unsigned char*, int, int&, int&, int&, int&) override { return (DECODING_STATE)7; }
  long SetOption (DECODER_OPTION, void*) override { return 8; }
  long GetOption (DECODER_OPTION, void*) override { return 9; }
};
struct bool_test_struct { char c; bool b; };

int vtable_check() {
  FakeDecoder f;
  ISVCDecoder* p = &f;
  int bad = lh264_check_decoder_vtable_c (p);
  if (bad) fprintf (stderr, "%d vtable entries are out of order between the C and C++ views\n", bad);
  if (lh264_c_bool_size() != sizeof (bool) || lh264_c_bool_offset() != offsetof (bool_test_struct, b)) { fprintf (stderr, "bool differs between C and C++\n"); bad++; }
  if (lh264_c_buffer_info_size() != sizeof (SBufferInfo) || lh264_c_decoding_param_size() != sizeof (SDecodingParam)) { fprintf (stderr, "struct sizes differ between C and C++\n"); bad++; }
  SDecoderCapability cap;                     // decoder_test.cpp:20-33
  if (WelsGetDecoderCapability (&cap) != 0 || cap.iProfileIdc != 66 || cap.iProfileIop != 0xE0 || cap.iLevelIdc != 32 || cap.iMaxMbps != 216000 || cap.iMaxFs != 5120 ||
      cap.iMaxCpb != 20000 || cap.iMaxDpb != 20480 || cap.iMaxBr != 20000 || cap.bRedPicCap != false) { fprintf (stderr, "WelsGetDecoderCapability differs from the reference's constants\n"); bad++; }
  return bad;
}

int decode_file (const char* path) {
  FILE* f = fopen (path, "rb");
  if (!f) { perror (path); return 2; }
  std::vector<unsigned char> data;
  unsigned char tmp[65536];
  size_t n;
  while ((n = fread (tmp, 1, sizeof (tmp), f)) > 0) data.insert (data.end(), tmp, tmp + n);
  fclose (f);
  ISVCDecoder* dec = nullptr;
  if (WelsCreateDecoder (&dec) != 0 || !dec) { fprintf (stderr, "WelsCreateDecoder failed\n"); return 3; }
  SDecodingParam prm;
  memset (&prm, 0, sizeof (prm));
  prm.eOutputColorFormat = videoFormatI420;
  prm.uiTargetDqLayer = UCHAR_MAX;
  prm.eEcActiveIdc = ERROR_CON_SLICE_COPY;
  prm.sVideoProperty.eVideoBsType = VIDEO_BITSTREAM_DEFAULT;
  if (dec->Initialize (&prm) != 0) { fprintf (stderr, "Initialize failed (no CUDA device? there is no CPU fallback)\n"); WelsDestroyDecoder (dec); return 4; }
  lh264host::Sha1 sha;
  int frames = 0, state_or = 0;
  auto one = [&] (const unsigned char* src, int len) {
    unsigned char* planes[3] = {nullptr, nullptr, nullptr};
    SBufferInfo bi;
    memset (&bi, 0, sizeof (bi));
    state_or |= (int)dec->DecodeFrame2 (src, len, planes, &bi);
    if (bi.iBufferStatus != 1) return;
    const int w = bi.UsrData.sSystemBuffer.iWidth, h = bi.UsrData.sSystemBuffer.iHeight;
    const int sy = bi.UsrData.sSystemBuffer.iStride[0], sc = bi.UsrData.sSystemBuffer.iStride[1];
    for (int r = 0; r < h; r++) sha.update (planes[0] + (size_t)r * sy, (size_t)w);
    for (int p = 1; p < 3; p++) for (int r = 0; r < h / 2; r++) sha.update (planes[p] + (size_t)r * sc, (size_t) (w / 2));
    frames++;
  };
  // ReadFrame: a unit runs from one 00 00 00 01 to the next (a unit never ends inside the first 4 bytes)
  size_t pos = 0;
  while (pos < data.size()) {
    size_t end = data.size();
    for (size_t i = pos + 4; i + 3 < data.size(); i++)
      if (data[i] == 0 && data[i + 1] == 0 && data[i + 2] == 0 && data[i + 3] == 1) { end = i; break; }
    one (data.data() + pos, (int) (end - pos));
    pos = end;
  }
  int eos = 1;
  dec->SetOption (DECODER_OPTION_END_OF_STREAM, &eos);
  one (nullptr, 0);
  dec->Uninitialize();
  WelsDestroyDecoder (dec);
  printf ("frames %d sha1 %s state_or %d\n", frames, sha.hex().c_str(), state_or);
  return 0;
}
} // namespace

int main (int argc, char** argv) {
  if (argc >= 2 && !strcmp (argv[1], "vtable")) return vtable_check();
  if (argc >= 2 && !strcmp (argv[1], "cinit")) { const long rc = lh264_c_create_init_destroy(); printf ("Initialize from C: %ld\n", rc); return rc == 0 ? 0 : 5; }
  if (argc >= 3 && !strcmp (argv[1], "decode")) return decode_file (argv[2]);
  fprintf (stderr, "usage: %s vtable | cinit | decode in.264\n", argv[0]);
  return 2;
}
