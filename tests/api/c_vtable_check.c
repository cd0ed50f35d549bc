/* c_vtable_check.c — compiled as C: the C view of ISVCDecoder in include/lh264_codec_api.h (a pointer to a table of function
 * pointers) must name the entries in the order the C++ class declares its virtual functions.  Same method as the reference's
 * test/api/c_interface_test.c (CheckDecoderInterface): every entry of a fake object returns its ordinal in the class declaration
 * (codec_api.h:353-453: Initialize, Uninitialize, DecodeFrame, DecodeFrameNoDelay, DecodeFrame2, DecodeParser, DecodeFrameEx, SetOption,
 * GetOption). */
#include "../../include/lh264_codec_api.h"
#include <stddef.h>

typedef int (*Func) (const void*);
#define CALL(p, m) (((Func)((*p)->m))(p))

int lh264_check_decoder_vtable_c (ISVCDecoder* p) {
  int bad = 0;
  bad += CALL (p, Initialize) != 1;
  bad += CALL (p, Uninitialize) != 2;
  bad += CALL (p, DecodeFrame) != 3;
  bad += CALL (p, DecodeFrameNoDelay) != 4;
  bad += CALL (p, DecodeFrame2) != 5;
  bad += CALL (p, DecodeParser) != 6;
  bad += CALL (p, DecodeFrameEx) != 7;
  bad += CALL (p, SetOption) != 8;
  bad += CALL (p, GetOption) != 9;
  return bad;
}

/* layout facts the reference checks between its C and C++ builds (c_interface_test.c:45-60) */
struct bool_test_struct { char c; bool b; };
size_t lh264_c_bool_size (void) { return sizeof (bool); }
size_t lh264_c_bool_offset (void) { return offsetof (struct bool_test_struct, b); }
size_t lh264_c_buffer_info_size (void) { return sizeof (SBufferInfo); }
size_t lh264_c_decoding_param_size (void) { return sizeof (SDecodingParam); }

/* a real decoder driven from C through the table (needs the library and a CUDA device): returns the Initialize code */
long lh264_c_create_init_destroy (void) {
  ISVCDecoder* dec = NULL;
  SDecodingParam prm;
  long rc;
  unsigned i;
  if (WelsCreateDecoder (&dec) != 0 || !dec) return -100;
  for (i = 0; i < sizeof (prm); i++) ((char*)&prm)[i] = 0;
  prm.eOutputColorFormat = videoFormatI420;
  prm.uiTargetDqLayer = 255;
  prm.eEcActiveIdc = ERROR_CON_SLICE_COPY;
  prm.sVideoProperty.eVideoBsType = VIDEO_BITSTREAM_DEFAULT;
  rc = (*dec)->Initialize (dec, &prm);
  (*dec)->Uninitialize (dec);
  WelsDestroyDecoder (dec);
  return rc;
}
