/* lh264_runtime.h — the multi-stream runtime behind the library boundary: Annex-B bytes of MANY independent
 * streams in host memory -> decoded I420 pictures in (pinned) host memory + the `.pip` file set of every stream.
 * Exported by libopenh264_lh264.so next to the OpenH264 decoder API (lh264_codec_api.h).
 *
 * What it replaces in the reference.  The reference decodes ONE stream per process, on one thread, with its model state
 * in process globals (codec/decoder/core/src/compression_stream.cpp:18-30): its driver is the unit loop of
 *   h264dec                  codec/console/dec/src/h264dec.cpp:246-364   (Annex-B unit -> DecodeFrameNoDelay -> write YUV)
 *   CWelsDecoder::DecodeFrame2 codec/decoder/plus/src/welsDecoderExt.cpp:408-470
 *   MultiFileWriter          codec/console/dec/src/h264dec.cpp:79-120    (x.pip + x.pip.<tag> at exit)
 * run once per input file.  This runtime is that loop for S streams at once: parse threads (one Decoder per stream, the
 * host CAVLC/CABAC parser of losslessh264_b200/host) write each picture's hand-off records straight into pinned staging
 * memory, a submit thread per pipeline hands every step's pictures (one per live stream) to the GPU with ONE
 * lh264_recon_submit + ONE lh264_model_submit (include/lh264_b200.h), decoded pictures come back with one packed copy per
 * step, and parsing of steps t+1 and t+2 overlaps the GPU work of step t.  Results are those of S independent h264dec runs:
 * the same I420 bytes (SHA-1 per stream) and the same `.pip` bytes (default stream + tag side streams).
 *
 * No CPU fallback: lh264_rt_run() fails with LH264_E_NO_DEVICE (-1) when there is no usable CUDA device.
 */
#ifndef LH264_RUNTIME_H_
#define LH264_RUNTIME_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lh264_rt lh264_rt;

typedef struct lh264_rt_config {
  int32_t device;        /* CUDA ordinal                                                                                 */
  int32_t threads;       /* parse threads over all pipelines (0 = hardware concurrency)                                  */
  int32_t pipelines;     /* independent pipelines (GPU context + submit thread + a share of the parse threads), >= 1     */
  int32_t want_yuv;      /* != 0: reconstruct and bring every decoded picture to the host (h264dec's YUV output)         */
  int32_t want_hash;     /* != 0: SHA-1 over the output pictures of every stream (needs want_yuv)                         */
  int32_t want_pip;      /* != 0: run the recompression model and build the `.pip` file set (h264dec in.264 out.pip)     */
  int32_t loops;         /* replay every input this many times as one long stream (>= 1; benchmarks)                      */
  int32_t max_pictures;  /* stop a stream after this many pictures (0 = no limit)                                         */
  int32_t depth;         /* staging buffers per parse thread = steps parsed ahead of the GPU + 1 (0 = default 3)          */
  int32_t reserved[7];
} lh264_rt_config;

/* One decoded picture handed to the caller: cropped I420, planes tightly packed (Y w x h, then U and V ((w+1)/2 x (h+1)/2)).
 * The memory is pinned staging owned by the runtime and valid until the callback returns.  Called from the pipeline's submit
 * thread, in decode order within a stream; pictures the reference would not output (before a stream's first IDR) are not
 * reported.                                                                                                               */
typedef void (*lh264_rt_picture_fn) (void* user, int stream, long picture_index, int width, int height, const uint8_t* i420);

typedef struct lh264_rt_stream_result {
  long    pictures;            /* pictures reconstructed / modelled                                                        */
  long    pictures_output;     /* of which handed out                                                                      */
  char    yuv_sha1[41];        /* hex SHA-1 over the output pictures (want_hash), else "-"                                  */
  char    pip_sha1[41];        /* hex SHA-1 over default stream, then (tag byte, tag stream) for every tag with a stream  */
  size_t  pip_bytes;           /* size of the whole `.pip` file set                                                        */
  int32_t error;               /* 0, or the last negative code the stream's decoder reported                               */
} lh264_rt_stream_result;

typedef struct lh264_rt_stats {
  long   streams, pictures;
  double seconds;              /* first parse to last picture back on the host (model side streams still on the device)    */
  double seconds_total;        /* including lh264_model_finish + read-back of every side stream                            */
  size_t input_bytes, pip_bytes;
  double submit_wait_parse_s;  /* submit threads: time spent waiting for the parse threads (summed over pipelines)         */
  double submit_wait_gpu_s;    /* ... waiting for the device (inputs consumed / pictures back)                             */
  double parse_busy_s;         /* parse threads: time spent parsing (summed over threads)                                  */
  double parse_wait_s;         /* ... waiting for a free staging buffer                                                    */
  unsigned long long kernel_launches;
  size_t h2d_bytes, d2h_bytes; /* bytes handed to the device / brought back, whole run                                     */
  double setup_s;              /* start of the run until step 0 (one picture per stream) has been handed to the device:
                                  opening every device stream and model stream happens here                                */
  long   setup_pictures;       /* pictures of step 0                                                                       */
  double pinned_alloc_s;       /* parse threads: time spent growing pinned staging memory (summed; part of parse_busy_s)   */
  double open_streams_s;       /* submit threads: time spent opening device streams and model streams                      */
  int32_t threads, pipelines;
} lh264_rt_stats;

/* cfg may be NULL (defaults: device 0, all cores, 1 pipeline, YUV + hash + pip).  Returns 0 or a negative LH264_E_* code.  */
int  lh264_rt_create (lh264_rt** out, const lh264_rt_config* cfg);
void lh264_rt_destroy (lh264_rt* rt);
/* Adds one stream: len bytes of Annex-B data.  The bytes are BORROWED and must stay valid and unchanged until lh264_rt_run()
 * returns; they must be followed by 4 readable bytes 00 00 00 01 when trailing_start_code != 0 (h264dec appends the same
 * sentinel, h264dec.cpp:238), otherwise the runtime makes its own copy with the sentinel appended.  Returns the stream
 * index (>= 0) or a negative error.                                                                                        */
int  lh264_rt_add_stream (lh264_rt* rt, const uint8_t* annexb, size_t len, int trailing_start_code);
/* Decodes every stream to its end.  fn may be NULL (pictures still travel to pinned host memory when want_yuv).            */
int  lh264_rt_run (lh264_rt* rt, lh264_rt_picture_fn fn, void* user);
int  lh264_rt_stream_result_get (const lh264_rt* rt, int stream, lh264_rt_stream_result* out);
/* Bytes of one file of a stream's `.pip` set after lh264_rt_run(): tag 0 = the default stream `x.pip`, tag t = `x.pip.<t>`.
 * Copies up to cap bytes into dst and returns the file's size (0 = the tag has no file), or a negative error.              */
long long lh264_rt_pip_read (const lh264_rt* rt, int stream, int tag, uint8_t* dst, size_t cap);
int  lh264_rt_stats_get (const lh264_rt* rt, lh264_rt_stats* out);
const char* lh264_rt_last_error (const lh264_rt* rt);
/* The pinned staging memory of finished runs is kept for the next run of the process (pinning is slow); this gives it back.   */
void lh264_rt_trim (void);

#ifdef __cplusplus
}
#endif
#endif /* LH264_RUNTIME_H_ */
