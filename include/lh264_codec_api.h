/*
 * lh264_codec_api.h — the decoder half of the OpenH264 v1.4 public API, as exported by
 * libopenh264_lh264.so.  ABI-compatible with the reference's codec/api/svc/codec_api.h:345-455
 * (ISVCDecoder virtual order), :482-520 (C vtable), :550-570 (factory functions) and with the
 * structs/enums of codec_app_def.h / codec_def.h it uses (same names, values and layouts), so a
 * caller built against the reference headers (h264dec.cpp:530-561, test/api/BaseDecoderTest.cpp)
 * links against this library unchanged.  Only decoder-side declarations are provided.
 */
#ifndef LH264_CODEC_API_H_
#define LH264_CODEC_API_H_

#include <stddef.h>
#ifdef __cplusplus
#define LH264_EXTAPI
#else
#include <stdbool.h>                      /* the C view (codec_api.h:482-520) uses bool in its structs, codec_app_def.h:43-47 */
#endif

/* codec_app_def.h:76-100 */
typedef enum {
  dsErrorFree = 0x00, dsFramePending = 0x01, dsRefLost = 0x02, dsBitstreamError = 0x04, dsDepLayerLost = 0x08,
  dsNoParamSets = 0x10, dsDataErrorConcealed = 0x20,
  dsInvalidArgument = 0x1000, dsInitialOptExpected = 0x2000, dsOutOfMemory = 0x4000, dsDstBufNeedExpan = 0x8000
} DECODING_STATE;

/* codec_app_def.h:150-166 */
typedef enum {
  DECODER_OPTION_DATAFORMAT = 0, DECODER_OPTION_END_OF_STREAM, DECODER_OPTION_VCL_NAL, DECODER_OPTION_TEMPORAL_ID,
  DECODER_OPTION_FRAME_NUM, DECODER_OPTION_IDR_PIC_ID, DECODER_OPTION_LTR_MARKING_FLAG, DECODER_OPTION_LTR_MARKED_FRAME_NUM,
  DECODER_OPTION_ERROR_CON_IDC, DECODER_OPTION_TRACE_LEVEL, DECODER_OPTION_TRACE_CALLBACK, DECODER_OPTION_TRACE_CALLBACK_CONTEXT,
  DECODER_OPTION_GET_STATISTICS,
  /* Extension of this library (not in OpenH264): SetOption (.., const char* path) switches the recompressor on, as the
   * reference's h264dec does when its OUTPUT name contains ".pip" (h264dec.cpp:167-169).  Units fed through
   * DecodeFrameNoDelay are then also modelled on the GPU; Uninitialize() (or a SetOption with NULL) writes `path` (the
   * default stream) and `path.<tag>` (one file per arithmetic-coded side stream), compression_stream.cpp:269-287.     */
  DECODER_OPTION_LH264_PIP_OUTPUT = 1000,
  /* A `path` ending in ".pipx" selects the single-file container, written WHILE the stream is compressed: every N units
   * (SetOption (DECODER_OPTION_LH264_PIP_FLUSH_UNITS, int* N), default 64, 0 = only at the end) the bytes of the default
   * stream and of every side stream that can no longer change are appended as chunks (format: host/pip_container.h) -
   * the reference keeps everything in memory until exit (FlushOnClose, h264dec.cpp:105-121).                           */
  DECODER_OPTION_LH264_PIP_FLUSH_UNITS = 1001
} DECODER_OPTION;

/* codec_app_def.h:171-180 */
typedef enum {
  ERROR_CON_DISABLE = 0, ERROR_CON_FRAME_COPY, ERROR_CON_SLICE_COPY, ERROR_CON_FRAME_COPY_CROSS_IDR, ERROR_CON_SLICE_COPY_CROSS_IDR,
  ERROR_CON_SLICE_COPY_CROSS_IDR_FREEZE_RES_CHANGE, ERROR_CON_SLICE_MV_COPY_CROSS_IDR, ERROR_CON_SLICE_MV_COPY_CROSS_IDR_FREEZE_RES_CHANGE
} ERROR_CON_IDC;

typedef enum { VIDEO_BITSTREAM_AVC = 0, VIDEO_BITSTREAM_SVC = 1, VIDEO_BITSTREAM_DEFAULT = VIDEO_BITSTREAM_SVC } VIDEO_BITSTREAM_TYPE;
typedef enum { videoFormatI420 = 23 } EVideoFormatType;     /* codec_def.h:56 — the only output format */
typedef enum { cmResultSuccess = 0, cmInitParaError, cmUnkonwReason, cmMallocMemeError, cmInitExpected, cmUnsupportedData } CM_RETURN;

typedef struct { unsigned int size; VIDEO_BITSTREAM_TYPE eVideoBsType; } SVideoProperty;
typedef struct TagSVCDecodingParam {     /* codec_app_def.h:485-496 */
  char* pFileNameRestructed;
  EVideoFormatType eOutputColorFormat;
  unsigned int uiCpuLoad;
  unsigned char uiTargetDqLayer;
  ERROR_CON_IDC eEcActiveIdc;
  bool bParseOnly;
  SVideoProperty sVideoProperty;
} SDecodingParam, *PDecodingParam;

typedef struct TagSysMemBuffer { int iWidth, iHeight, iFormat; int iStride[2]; } SSysMEMBuffer;   /* codec_def.h:187-192 */
typedef struct TagBufferInfo {           /* codec_def.h:197-204 */
  int iBufferStatus;
  unsigned long long uiInBsTimeStamp, uiOutYuvTimeStamp;
  union { SSysMEMBuffer sSystemBuffer; } UsrData;
} SBufferInfo;
typedef enum { FEEDBACK_NON_VCL_NAL = 0, FEEDBACK_VCL_NAL, FEEDBACK_UNKNOWN_NAL } FEEDBACK_VCL_NAL_IN_AU;   /* codec_app_def.h:184-188 */
typedef struct TagDecoderCapability {    /* codec_app_def.h:590-600 */
  int iProfileIdc, iProfileIop, iLevelIdc, iMaxMbps, iMaxFs, iMaxCpb, iMaxDpb, iMaxBr;
  bool bRedPicCap;
} SDecoderCapability;
typedef struct _tagVersion { unsigned int uMajor, uMinor, uRevision, uReserved; } OpenH264Version;
typedef struct TagParserBsInfo SParserBsInfo;   /* DecodeParser: parse-only mode is not provided */
typedef struct TagDecoderStatistics {    /* codec_app_def.h: SDecoderStatistics */
  unsigned int uiWidth, uiHeight;
  float fAverageFrameSpeedInMs, fActualAverageFrameSpeedInMs;
  unsigned int uiDecodedFrameCount, uiResolutionChangeTimes, uiIDRCorrectNum, uiAvgEcRatio, uiAvgEcPropRatio, uiEcIDRNum, uiEcFrameNum,
           uiIDRLostNum, uiFreezingIDRNum, uiFreezingNonIDRNum;
  int iAvgLumaQp;
} SDecoderStatistics;

#ifdef __cplusplus
class ISVCDecoder {                       /* virtual order: codec_api.h:345-455 */
 public:
  virtual long Initialize (const SDecodingParam* pParam) = 0;
  virtual long Uninitialize() = 0;
  virtual DECODING_STATE DecodeFrame (const unsigned char* pSrc, const int iSrcLen, unsigned char** ppDst, int* pStride, int& iWidth, int& iHeight) = 0;
  virtual DECODING_STATE DecodeFrameNoDelay (const unsigned char* pSrc, const int iSrcLen, unsigned char** ppDst, SBufferInfo* pDstInfo) = 0;
  virtual DECODING_STATE DecodeFrame2 (const unsigned char* pSrc, const int iSrcLen, unsigned char** ppDst, SBufferInfo* pDstInfo) = 0;
  virtual DECODING_STATE DecodeParser (const unsigned char* pSrc, const int iSrcLen, SParserBsInfo* pDstInfo) = 0;
  virtual DECODING_STATE DecodeFrameEx (const unsigned char* pSrc, const int iSrcLen, unsigned char* pDst, int iDstStride, int& iDstLen,
                                        int& iWidth, int& iHeight, int& iColorFormat) = 0;
  virtual long SetOption (DECODER_OPTION eOptionId, void* pOption) = 0;
  virtual long GetOption (DECODER_OPTION eOptionId, void* pOption) = 0;
  virtual ~ISVCDecoder() {}
};
extern "C" {
#else
typedef struct ISVCDecoderVtbl ISVCDecoderVtbl;
typedef const ISVCDecoderVtbl* ISVCDecoder;
struct ISVCDecoderVtbl {                  /* codec_api.h:482-520 */
  long (*Initialize) (ISVCDecoder*, const SDecodingParam* pParam);
  long (*Uninitialize) (ISVCDecoder*);
  DECODING_STATE (*DecodeFrame) (ISVCDecoder*, const unsigned char* pSrc, const int iSrcLen, unsigned char** ppDst, int* pStride, int* iWidth, int* iHeight);
  DECODING_STATE (*DecodeFrameNoDelay) (ISVCDecoder*, const unsigned char* pSrc, const int iSrcLen, unsigned char** ppDst, SBufferInfo* pDstInfo);
  DECODING_STATE (*DecodeFrame2) (ISVCDecoder*, const unsigned char* pSrc, const int iSrcLen, unsigned char** ppDst, SBufferInfo* pDstInfo);
  DECODING_STATE (*DecodeParser) (ISVCDecoder*, const unsigned char* pSrc, const int iSrcLen, SParserBsInfo* pDstInfo);
  DECODING_STATE (*DecodeFrameEx) (ISVCDecoder*, const unsigned char* pSrc, const int iSrcLen, unsigned char* pDst, int iDstStride, int* iDstLen,
                                   int* iWidth, int* iHeight, int* iColorFormat);
  long (*SetOption) (ISVCDecoder*, DECODER_OPTION eOptionId, void* pOption);
  long (*GetOption) (ISVCDecoder*, DECODER_OPTION eOptionId, void* pOption);
};
#endif

/* welsDecoderExt.cpp:666-708, codec_api.h:550-570 */
long WelsCreateDecoder (ISVCDecoder** ppDecoder);
void WelsDestroyDecoder (ISVCDecoder* pDecoder);
int WelsGetDecoderCapability (SDecoderCapability* pDecCapability);
OpenH264Version WelsGetCodecVersion (void);
void WelsGetCodecVersionEx (OpenH264Version* pVersion);

/* ---- Extension of this library (not in OpenH264): the `.pip` container in the decompress direction -------------------
 * What the reference's h264dec does around the model when its INPUT name contains ".pip" (h264dec.cpp:170-173), split at
 * the hot-path boundary so that many files can share one lh264_model_decode() call (include/lh264_b200.h):
 *   plan     = lh264_pip_plan (default stream)            parameter sets + slice headers -> lh264_decode_slice list
 *   lh264_model_decode (ctx, jobs ...)                    side streams -> slice_data() bits, on the GPU
 *   lh264_pip_assemble (plan, default stream, bits ...)   -> the original Annex-B file                                  */
struct lh264_decode_slice;
struct lh264_decode_slice_out;
typedef struct lh264_pip_plan_s lh264_pip_plan_t;
/* NULL on failure, with the reason in err (CABAC slices, slices dropped at compression time ...) */
lh264_pip_plan_t* lh264_pip_plan (const unsigned char* default_stream, size_t len, char* err, size_t err_cap);
void lh264_pip_plan_free (lh264_pip_plan_t* plan);
/* slices of the plan (n_slices of them), and what lh264_decode_job needs: max_mbs, with_8x8 */
const struct lh264_decode_slice* lh264_pip_plan_slices (const lh264_pip_plan_t* plan, int* n_slices, int* max_mbs, int* with_8x8);
/* returns the length of the expanded file (copied into dst when it fits in cap), or -1 */
long long lh264_pip_assemble (const lh264_pip_plan_t* plan, const unsigned char* default_stream, const unsigned char* slice_bits,
                              const struct lh264_decode_slice_out* outs, unsigned char* dst, size_t cap);

/* Single-file container (an addition, the reference only knows the file set): `x.pipx` holds the default stream
 * and every tag stream of a `.pip` file set behind a table of (tag, length) entries (format: host/pip_container.h).  Lossless in
 * both directions; 0 on success, -1 cannot read the input, -2 cannot write the output.  h264dec_lh264 accepts `.pipx` wherever
 * the reference CLI takes `.pip`.                                                                                              */
int lh264_pip_pack (const char* pip_path, const char* archive_path);
int lh264_pip_unpack (const char* archive_path, const char* pip_path);

#ifdef __cplusplus
}
#endif
#endif
